"""Pretty-print gpurun_out/kernels_detail.json (written by bench.py): per (kernel, shape) CUDA-event times."""
import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/kernels_detail.json'))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
print('total kernel ms/step', round(sum(x['ms_per_step'] for x in d), 2))
for x in d[:n]:
    print(f"{x['kernel']:22s} {x['tag']:28s} n={x['launches_per_step']:5.1f} ms={x['ms_per_step']:6.2f} us/l={x['us_per_launch']:8.1f} TF={x['tflops']:6.1f} GB/s={x['gbs']:7.1f}")
fam = {}
for x in d:
    fam[x['kernel']] = fam.get(x['kernel'], 0) + x['ms_per_step']
print({k: round(v, 2) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])})
