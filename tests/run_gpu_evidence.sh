#!/bin/bash
# One GPU visit that regenerates the committed evidence of a build: full GPU test suite, smoke, the bench lines (C2 headline, C3, C5),
# the ncu launch list of bench.py, and the ncu --set full capture of every hot kernel (exported as a raw-page CSV: the report itself
# exceeds the 64 MB that travel back).  Afterwards, here:  python profiles/ncu_summary.py gpurun_out/r2_prof_all_raw.csv
#   profiles/r2_ncu_summary.csv profiles/ncu_traffic.json gpurun_out/prof_kernels_all.json profiles/r2_ncu_metrics.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  OT_LOGIT_TOL=1e-2 OT_GRAD_TOL=3e-2 timeout 900 python -m pytest tests -q -m gpu -x --timeout 180 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_smoke.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r2_bench_c2.err | tail -1 > gpurun_out/r2_bench_c2.json && cp gpurun_out/kernels_detail.json gpurun_out/r2_kernels_detail.json
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c2.json').read()); print('c2', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])"
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 2>gpurun_out/r2_bench_c3.err | tail -1 > gpurun_out/r2_bench_c3_1gpu.json
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c3_1gpu.json').read()); print('c3', d['ms_per_step'], d['value'])"
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 2>gpurun_out/r2_bench_c5.err | tail -1 > gpurun_out/r2_bench_c5.json
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c5.json').read()); print('c5', d['ms_per_step'], d['value'], d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_raw.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2_launches_ncu.log 2>&1
PROF_ONCE=1 timeout 200 python profiles/prof_kernels.py all > gpurun_out/r2_prof_all_plain.log 2>&1
PROF_ONCE=1 timeout 900 ncu --set full --clock-control none --import-source on -o /tmp/r2_prof_all -f python profiles/prof_kernels.py all > gpurun_out/r2_prof_all_ncu.log 2>&1
ncu -i /tmp/r2_prof_all.ncu-rep --page raw --csv > gpurun_out/r2_prof_all_raw.csv 2>/dev/null
ls -la /tmp/r2_prof_all.ncu-rep gpurun_out/r2_prof_all_raw.csv gpurun_out/r2_launches_raw.csv | awk '{print $5, $9}'
