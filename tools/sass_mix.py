"""Instruction mix (executed warp instructions per opcode) from an `ncu --page source --csv` export."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]; si = hdr.index('Source'); ii = hdr.index('Instructions Executed')
c = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    if len(r) <= ii or not r[ii].isdigit(): continue
    s = r[si].split()
    op = s[1] if s[0].startswith('@') else s[0]
    p = op.rstrip(';').split('.')
    op = p[0] + ('.' + p[1] if len(p) > 1 and p[0].startswith(('F', 'I2F', 'F2I', 'LD', 'ST', 'IMAD', 'VI')) else '')
    n = int(r[ii]); c[op] += n; tot += n
print("total warp instructions", tot)
for k, v in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30): print(f"{k:20s} {v:12d} {100*v/tot:6.2f}%")
