"""Reduce `ncu --page source --csv` output of one kernel to (a) the dynamic opcode mix and (b) contiguous SASS regions
with similar execution counts (roughly: one line per loop body), with their share of executed instructions and of
stall samples.   usage: ncu -i rep --page source --csv --kernel-id :::N > src.csv ; python profiles/ncu_regions.py src.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 10 and r[0].startswith('0x')]
data = data[:len(data) // 2] if len(data) > 1 and data[0][0] == data[len(data) // 2][0] else data   # page is emitted twice
iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iE]) for r in data)
tots = sum(int(r[iN]) for r in data)
print(rows[0][1][:100])
print('warp instructions executed', tot, ' stall samples', tots)
byop, sop = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iS].strip())
    op = m.group(2) if m else r[iS][:8]
    byop[op] += int(r[iE]); sop[op] += int(r[iN])
print('-- opcode mix')
for op, c in byop.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 18):
    print(f'{op:10s} {100 * c / tot:5.1f}% of instr  {100 * sop[op] / max(tots, 1):5.1f}% of samples')
print('-- regions')
prev = None; start = 0; acc = 0; n = 0; samp = 0
for i, r in enumerate(data + [None]):
    e = int(r[iE]) if r else -1
    if prev is None or r is None or abs(e - prev) > 0.2 * max(e, prev, 1):
        if prev is not None and (acc > tot * 0.01 or samp > tots * 0.02):
            print(f'sass {start:5d}-{i - 1:5d} n={n:4d} exec/instr {prev:10d} instr {100 * acc / tot:5.1f}%  samples {100 * samp / max(tots, 1):5.1f}%   {data[start][iS].strip()[:48]}')
        start = i; acc = 0; n = 0; samp = 0
    if r is None:
        break
    prev = e; acc += e; n += 1; samp += int(r[iN])
