"""Development aid: per-instruction stall samples of an `ncu --page source --csv` export (top N SASS lines by samples)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 45
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
keys = ['stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_branch_resolving', 'stall_lg', 'stall_math', 'stall_no_inst', 'stall_not_selected', 'stall_barrier', 'stall_membar', 'stall_sleep', 'stall_mio']
print('total samples', sum(f(r, '# Samples') for r in data), {k: sum(f(r, k) for r in data) for k in keys}, 'insts', sum(f(r, 'Instructions Executed') for r in data))
top = sorted(range(len(data)), key=lambda i: -f(data[i], '# Samples'))[:n]
for i in sorted(top):
    r = data[i]
    print(i, r[ix['Source']][:72].ljust(72), int(f(r, '# Samples')), ' '.join('%s=%d' % (k[6:], f(r, k)) for k in keys if f(r, k) > 0), 'exec', int(f(r, 'Instructions Executed')))
