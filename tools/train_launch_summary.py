"""Summarise an ncu launch list of the train step (last step of the run) by kernel: python tools/train_launch_summary.py CSV [steps]"""
import csv, collections, re, sys
fn = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else None
lines = [l for l in open(fn) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
def us(r):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    return v / 1000 if u.startswith('n') else v if u.startswith('u') else v * 1000
# one step = from one pack_multi_kernel (first launch of the forward) to the next
starts = [i for i, r in enumerate(rows) if 'pack_multi_kernel' in r['Kernel Name']]
if len(starts) >= 2:
    rows = rows[starts[-2]:starts[-1]]
elif starts:
    rows = rows[starts[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for r in rows:
    k = re.sub(r'\(.*', '', r['Kernel Name']); k = re.sub(r'void |<unnamed>::|at::native::|at::', '', k)[:64]
    agg[k][0] += 1; agg[k][1] += us(r); tot += us(r)
print(f"{fn}: {len(rows)} launches, {tot/1000:.2f} ms kernel time in one step")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
    print(f"| `{k}` | {c} | {t:.1f} | {100*t/tot:.1f}% |")
