"""Summarise an `ncu --page source --csv` export: executed-instruction histogram by opcode and the
hottest address ranges.  usage: sass_hot.py file.csv [min_exec]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iE]) for r in data)
print("total warp instr", tot, "static", len(data))
mn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ops = collections.Counter(); samp = collections.Counter()
for r in data:
    e = int(r[iE])
    if e < mn: continue
    s = r[iS].strip()
    if s.startswith("@"): s = s.split(None, 1)[1]
    op = s.split()[0].split(".")[0]
    ops[op] += e; samp[op] += int(r[iSamp])
tots = sum(samp.values())
for op, e in ops.most_common(40):
    print(f"{op:10s} {e:12d} {100*e/tot:6.2f}%  samples {100*samp[op]/max(tots,1):6.2f}%")
