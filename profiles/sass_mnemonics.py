"""Static evidence from the built library (no GPU needed): per kernel, how often the SASS mnemonics that prove the Blackwell
path appear (B200_PROFILING.md: UTCHMMA/UTCQMMA = tcgen05.mma, LDTM/STTM = TMEM access, UTMALDG/UTMASTG = TMA tensor copies,
SYNCS = mbarrier, REDG/ATOMG = global reductions, FFMA2 = packed fp32 math, MUFU = special-function unit, LDG.E.128 = 16-byte loads).
usage: python profiles/sass_mnemonics.py > profiles/r1_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'recommend_b200', 'lib', 'libonetrans_sm100.so')
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTMAPF', 'SYNCS', 'REDG', 'RED.', 'ATOMG', 'ATOMS', 'FFMA2', 'FMUL2',
         'FADD2', 'MUFU.TANH', 'MUFU.EX2', 'MUFU.LG2', 'MUFU.RSQ', 'LDG.E.128', 'STG.E.128', 'LDS.128', 'STS.128', 'HMMA', 'LDL', 'STL']
sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
kernel, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kernel = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
        kernel = re.sub(r'^void ', '', kernel)
        counts[kernel] = collections.Counter()
        total[kernel] = 0
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and kernel:
        total[kernel] += 1
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w) or (w.endswith('.') and op.startswith(w[:-1] + '.')) or ('.' in w and w in op):
                counts[kernel][w] += 1
                break
print(f'# {os.path.relpath(LIB)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a); local-memory LDL/STL = spills')
for k, c in counts.items():
    items = ' '.join(f'{w}={c[w]}' for w in WATCH if c[w])
    print(f'{k:70s} insts={total[k]:6d}  {items}')
