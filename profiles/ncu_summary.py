"""Reduce an `ncu --set full` report to the handful of metrics DESIGN.md / README.md quote, one row per launch, and
write the per-kernel DRAM traffic table bench.py reads for `roofline.traffic`.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_summary.csv [profiles/ncu_traffic.json
       gpurun_out/prof_kernels_all.json]   (the last file is prof_kernels.py's launch list: shape tags + algorithmic bytes)"""
import csv, json, subprocess, sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
           'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
           'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'inst_executed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']
FAMILY = {'ot_mixed_gemm_kernel': 'ot_mixed_gemm', 'ot_ffn_fused_kernel': 'ot_ffn_fwd', 'ot_wgrad_kernel': 'ot_wgrad', 'ot_attn_fwd_ws_kernel': 'ot_attn_fwd',
          'ot_attn_fwd_v2_kernel': 'ot_attn_fwd', 'ot_attn_fwd_v3_kernel': 'ot_attn_fwd', 'ot_attn_fwd_v4_kernel': 'ot_attn_fwd', 'ot_attn_fwd_v5_kernel': 'ot_attn_fwd', 'ot_attn_bwd_v2_kernel': 'ot_attn_bwd', 'ot_attn_cached_kernel': 'ot_attn_ns_cached_fwd', 'ot_attn_dkv_kernel': 'ot_attn_bwd', 'ot_attn_dq_kernel': 'ot_attn_bwd',
          'ot_attn_fwd_kernel': 'ot_attn_fwd', 'ot_attn_bwd_fused_kernel': 'ot_attn_bwd', 'rmsnorm_fwd_kernel': 'ot_rmsnorm_fwd',
          'rmsnorm_bwd_kernel': 'ot_rmsnorm_bwd'}

rep, out_csv = sys.argv[1], sys.argv[2]
# `rep` is an .ncu-rep, or the `ncu -i rep --page raw --csv` export of one (reports over 64 MB do not travel back from the GPU box)
raw = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [m for m in METRICS if m in hdr]
traffic = {}
with open(out_csv, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['id', 'kernel'] + [f'{m} [{units[hdr.index(m)]}]' for m in cols])
    for r in data:
        name = r[hdr.index('Kernel Name')]
        short = name.split('(')[0].replace('void ', '').replace('ot::', '')
        w.writerow([r[0], short] + [r[hdr.index(m)] for m in cols])
        fam = next((v for k, v in FAMILY.items() if k in name), None)
        if fam:
            def gb(m):
                v, u = float(r[hdr.index(m)].replace(',', '')), units[hdr.index(m)]
                return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}[u]
            t = gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum')
            traffic.setdefault(fam, []).append({'kernel': short, 'dram_bytes': t,
                                                'time_us': float(r[hdr.index('gpu__time_duration.sum')].replace(',', '')) *
                                                {'ms': 1e3, 'us': 1.0, 'ns': 1e-3, 's': 1e6}[units[hdr.index('gpu__time_duration.sum')]]})
print(f'wrote {out_csv}: {len(data)} launches, {len(cols)} metrics')
if len(sys.argv) > 3:
    out = {}
    if len(sys.argv) > 4:
        # join with prof_kernels.py's own launch list (same order per family): traffic / algorithmic bytes per shape
        launches = json.load(open(sys.argv[4]))
        per_fam = {}
        for l in launches:
            per_fam.setdefault(l['kernel'], []).append(l)
        for fam, v in traffic.items():
            ls = per_fam.get(fam, [])
            if len(ls) != len(v):
                continue
            for l, m in zip(ls, v):
                m['tag'], m['algorithmic_bytes'] = l['tag'], l['bytes']
                out[f"{fam}[{l['tag']}]"] = {'dram_bytes': m['dram_bytes'], 'algorithmic_bytes': l['bytes'],
                                             'traffic_over_algorithmic': m['dram_bytes'] / l['bytes'] if l['bytes'] else None}
    for fam, v in traffic.items():   # one number per family as well: the launch with the largest traffic
        out[fam] = max(v, key=lambda d: d['dram_bytes'])['dram_bytes']
    out['_detail'] = traffic
    out['_note'] = 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, layer-0 shapes of C2 (profiles/prof_kernels.py)'
    json.dump(out, open(sys.argv[3], 'w'), indent=1)
    print('wrote', sys.argv[3])

if len(sys.argv) > 5:
    tp = 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'
    def row_of(kname):
        return next(((i, r) for i, r in enumerate(data) if kname in r[hdr.index('Kernel Name')]), (None, None))
    fwd_name = next((k for k in ('ot_attn_fwd_v5_kernel', 'ot_attn_fwd_v3_kernel') if row_of(k)[1]), 'ot_attn_fwd_v3_kernel')   # layer 0: v5 since the tile-count dispatch
    (i_f, r_f), (i_b, r_b) = row_of(fwd_name), row_of('ot_attn_bwd_v2_kernel')
    if r_f and r_b and tp in hdr:
        def ms(r):
            return float(r[hdr.index('gpu__time_duration.sum')].replace(',', '')) * {'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}[units[hdr.index('gpu__time_duration.sum')]]
        f, b = float(r_f[hdr.index(tp)]), float(r_b[hdr.index(tp)])
        json.dump({'attn_tensor_pipe_util_pct': {'ot_attn_fwd[Lq458_Lk544]': f, 'ot_attn_bwd[Lq458_Lk544]': b,
                                                 'time_weighted': (f * ms(r_f) + b * ms(r_b)) / (ms(r_f) + ms(r_b)), 'metric': tp,
                                                 'source': f'{out_csv} rows {i_f} and {i_b} ({fwd_name}, ot_attn_bwd_v2_kernel; C2 layer 0: B 2048, H 4, '
                                                           'Lq 458, Lk 544, head_dim 64), same build as the bench'}},
                  open(sys.argv[5], 'w'), indent=1)
        print('wrote', sys.argv[5])

# ---- ordered join: the i-th library call of prof_kernels.py <-> its main kernel row of the capture (the delta pass of the attention
# backward and the second kernel of the two-kernel head_dim-96 backward are attributed to the same call) ----
if len(sys.argv) > 4:
    MAIN = [('ot_mixed_gemm_kernel', ('ot_mixed_gemm',)), ('ot_ffn_fused_kernel', ('ot_ffn_fwd', 'ot_ffn_bwd')), ('ot_wgrad_kernel', ('ot_wgrad',)),
            ('ot_attn_fwd_v5_kernel', ('ot_attn_fwd',)), ('ot_attn_fwd_v4_kernel', ('ot_attn_fwd',)), ('ot_attn_fwd_v3_kernel', ('ot_attn_fwd',)), ('ot_attn_fwd_ws_kernel', ('ot_attn_fwd',)), ('ot_attn_fwd_v2_kernel', ('ot_attn_fwd',)),
            ('ot_attn_fwd_kernel', ('ot_attn_fwd',)), ('ot_attn_bwd_v2_kernel', ('ot_attn_bwd',)), ('ot_attn_bwd_fused_kernel', ('ot_attn_bwd',)),
            ('ot_attn_dkv_kernel', ('ot_attn_bwd',)), ('rmsnorm_bwd_kernel', ('ot_rmsnorm_bwd',)), ('rmsnorm_fwd_kernel', ('ot_rmsnorm_fwd',)),
            ('ot_attn_cached_kernel', ('ot_attn_ns_cached_fwd',))]
    launches = json.load(open(sys.argv[4]))
    out = json.load(open(sys.argv[3]))
    def num(r, m):
        return float(r[hdr.index(m)].replace(',', '')) if m in hdr and r[hdr.index(m)] not in ('', 'n/a') else None
    def gbv(r, m):
        v, u = float(r[hdr.index(m)].replace(',', '')), units[hdr.index(m)]
        return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}[u]
    li = 0
    for i, r in enumerate(data):
        name = r[hdr.index('Kernel Name')]
        fams = next((f for k, f in MAIN if k in name), None)
        if fams is None:
            continue
        while li < len(launches) and launches[li]['kernel'] not in fams:
            li += 1
        if li == len(launches):
            break
        l = launches[li]
        li += 1
        t = gbv(r, 'dram__bytes_read.sum') + gbv(r, 'dram__bytes_write.sum')
        out[f"{l['kernel']}[{l['tag']}]"] = {
            'dram_bytes': t, 'algorithmic_bytes': l['bytes'], 'traffic_over_algorithmic': t / l['bytes'] if l['bytes'] else None,
            'capture': f'{out_csv} row {i} (ncu --set full --clock-control none, profiles/prof_kernels.py all, PROF_ONCE=1; main kernel of the call)',
            'time_ms': num(r, 'gpu__time_duration.sum') * {'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}[units[hdr.index('gpu__time_duration.sum')]],
            'tensor_pipe_pct': num(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
            'xu_pipe_pct': num(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
            'issue_active_pct': num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
            'lts_throughput_pct': num(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
            'l2_to_sm_read_bytes': gbv(r, 'l1tex__m_xbar2l1tex_read_bytes.sum') if 'l1tex__m_xbar2l1tex_read_bytes.sum' in hdr else None}
    json.dump(out, open(sys.argv[3], 'w'), indent=1)
    print('joined', li, 'of', len(launches), 'library calls')
