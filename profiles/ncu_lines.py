"""Top CUDA source lines of an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump
(stall samples, instructions, shared-memory wavefronts).  usage: ncu_lines.py dump.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No')
hdr = rows[hi]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
def num(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0
lines = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].strip().isdigit()]
total = sum(num(r, '# Samples') for r in lines)
print(f'total samples {total:.0f}')
cols = ('stall_long_sb', 'stall_short_sb', 'stall_barrier', 'stall_mio', 'stall_math', 'stall_wait', 'stall_lg', 'stall_branch_resolving')
print(f'{"line":>5} {"samp%":>6} {"inst":>10} {"shwave":>10} {"ideal":>10}  long short  barr   mio  math  wait    lg  brch | source')
for r in sorted(lines, key=lambda r: -num(r, '# Samples'))[:top]:
    print(f'{r[0]:>5} {100*num(r,"# Samples")/max(total,1):6.2f} {num(r,"Instructions Executed"):10.0f} {num(r,"L1 Wavefronts Shared"):10.0f} {num(r,"L1 Wavefronts Shared Ideal"):10.0f} '
          + ' '.join(f'{num(r,c):5.0f}' for c in cols) + f' | {r[1].strip()[:100]}')
