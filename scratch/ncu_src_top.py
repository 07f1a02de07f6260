"""Top stall lines of an `ncu --page source --csv` dump: python scratch/ncu_src_top.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = []
hdr = None
for r in rows:
    if len(r) > 4 and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    try:
        samp = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    ex = r[hdr.index("Instructions Executed")]
    stalls = {h: int(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit() and int(r[i]) > 0}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    out.append((samp, r[0], r[1][:90], ex, top))
tot = sum(o[0] for o in out)
print("total samples", tot)
for o in sorted(out, key=lambda o: -o[0])[:n]:
    print(f"{o[0]:6d} {100*o[0]/max(tot,1):5.1f}% ex={o[3]:>8s} {o[2]:90s} {o[4]}")
