import csv,sys
rows=list(csv.reader(sys.stdin))
blocks=[]; cur=None
for r in rows:
    if r and r[0]=='Kernel Name': cur=[]; blocks.append(cur); continue
    if cur is not None: cur.append(r)
b=blocks[int(sys.argv[1]) if len(sys.argv)>1 else 0]; hdr=b[0]; data=b[1:]
si=hdr.index('# Samples'); ii=hdr.index('Instructions Executed')
acc=0; accI=0
for k,r in enumerate(data):
    op=r[1].strip()
    acc+=int(r[si]); accI+=int(r[ii])
    if ('BAR.SYNC' in op or 'WARPSYNC' in op or 'REDUX' in op or 'EXIT' in op):
        if acc>0: print(f'{k:5d} samples={acc:5d} instr={accI:9d}  upto: {op[:50]}  (exec {r[ii]})')
        acc=0; accI=0
print('tail',acc,accI)
