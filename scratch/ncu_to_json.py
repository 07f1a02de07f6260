"""profiles/r2_k1_ncu_full.json from `ncu --set full` captures of k_sae_update_ts:
python scratch/ncu_to_json.py key=file.ncu-rep[:launch_index] ... > profiles/r2_k1_ncu_full.json"""
import csv, io, json, subprocess, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
        "sm__cycles_active.avg", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for arg in sys.argv[1:]:
    key, path = arg.split("=", 1)
    idx = 0
    if ":" in path:
        path, idx = path.rsplit(":", 1)
        idx = int(idx)
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units, r = rows[0], rows[1], rows[2 + idx]
    m = {k: [float(r[h.index(k)].replace(",", "")), units[h.index(k)]] for k in KEEP if k in h}
    rd = m["dram__bytes_read.sum"][0] * SCALE[m["dram__bytes_read.sum"][1]]
    wr = m["dram__bytes_write.sum"][0] * SCALE[m["dram__bytes_write.sum"][1]]
    out[key] = {"kernel": "k_sae_update_ts", "traffic_bytes_per_launch": int(rd + wr), "dram_read_bytes": int(rd),
                "dram_write_bytes": int(wr), "ncu_duration_us": m["gpu__time_duration.sum"][0], "metrics": m,
                "how": "ncu --set full --clock-control none --import-source on -k regex:k_sae_update_ts (default "
                       "cache control: caches flushed before every replay, so every state tile comes from HBM)"}
print(json.dumps(out, indent=1))
