"""Trim an `ncu --metrics gpu__time_duration.sum --csv` log of `bench.py` to (id, kernel, grid, block, ns) and print the
per-kernel share table.   usage: python profiles/launch_list.py gpurun_out/r1_launches_raw.csv profiles/r1_launches.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mu, g, b = (hdr.index(x) for x in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Grid Size', 'Block Size'))


def short(n):
    n = re.sub(r'\(.*', '', n).replace('void ', '').replace('ot::', '')
    return re.sub(r'<.*', '', n) if n.startswith('at::') or 'cutlass' in n or 'cublas' in n else n


# keep whole steps only: every step starts with the memset of the flat fp32 gradient buffer (FillFunctor<float>, > 100k blocks)
starts = [i for i, r in enumerate(data) if 'FillFunctor<float>' in r[kn] and int(r[g].strip('()').split(',')[0]) > 100000]
starts = [s0 for j, s0 in enumerate(starts) if j + 1 == len(starts) or starts[j + 1] - s0 > 10]   # (the buffer's own zero-init)
if len(starts) >= 2:
    data = data[starts[0]:starts[-1]]
    print(f'{len(starts) - 1} complete step(s): launches {starts[0]}..{starts[-1] - 1}')
agg, tot = collections.OrderedDict(), 0.0
with open(sys.argv[2], 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['id', 'kernel', 'grid', 'block', 'gpu__time_duration.sum [ns]'])
    for r in data:
        t = float(r[mv].replace(',', '')) * {'ns': 1.0, 'us': 1e3, 'ms': 1e6}[r[mu]]
        k = short(r[kn])
        w.writerow([r[0], k, r[g], r[b], int(t)])
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
print(f'{len(data)} launches, {tot / 1e6:.2f} ms serialised')
print('| kernel | launches | ms | share |\n|---|---|---|---|')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f'| `{k}` | {n} | {t / 1e6:.2f} | {100 * t / tot:.1f} % |')
