"""Summarise an `ncu --page source --csv` dump: top instructions by stall samples.
usage: python profiles/srcstalls.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = []
hdr = None
for r in rows:
    if len(r) > 3 and r[0] == "Address":
        hdr = r
        si = hdr.index('Source'); ns = hdr.index('# Samples')
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None or len(r) <= ns or not r[ns].isdigit():
        if len(r) >= 2 and r[0] == "Kernel Name":
            out.append(("K", r[1][:100]))
        continue
    out.append(("I", int(r[ns]), r[si].strip(), {h[6:]: int(r[i]) for i, h in stall_cols if r[i].isdigit() and int(r[i]) > 0}))
cur = []
def flush():
    if cur:
        tot = sum(x[1] for x in cur)
        print("  total samples", tot)
        for x in sorted(cur, key=lambda x: -x[1])[:N]:
            print(f"  {x[1]:6d} {100*x[1]/max(tot,1):5.1f}%  {x[2][:64]:64s} {x[3]}")
for x in out:
    if x[0] == "K":
        flush(); cur = []
        print(x[1])
    else:
        cur.append(x)
flush()
