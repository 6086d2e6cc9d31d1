"""CPU tests of the host logic: configuration mirror, pyramid scheduler (bit-exact index selection),
weight-group segments, and the C-ABI library (loads, exports every declared symbol, struct layouts match
the ctypes mirror).  No kernel is launched here."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from oracle import onetrans_oracle as O
import recommend_b200 as R
from recommend_b200 import _lib, ops
from recommend_b200.schedule import (keep_lens_halving, keep_lens_linear_to_ns, keep_lens_reference_ratio)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'onetrans_b200.h')


def test_config_mirrors_reference_defaults():
    c = R.OneTransConfig()   # OT/config.py:14-69
    assert (c.hidden_dim, c.num_layers, c.num_heads, c.ffn_dim) == (384, 8, 4, 1536)
    assert c.num_ns_tokens == 12 and c.max_seq_len == 2048 and c.dropout_rate == 0.1
    assert c.pyramid_ratios == [0.5, 0.3, 0.2, 0.1, 0.05, 0.03, 0.02, 0.01]
    assert c.feature_config['sequence_features'] == ['click_seq', 'cart_seq', 'purchase_seq']
    assert c.tasks == ['ctr', 'cvr'] and c.gradient_clip_norm == 90.0
    s = R.get_model_config('small')
    assert (s.hidden_dim, s.num_layers, s.ffn_dim, s.num_heads) == (256, 6, 1024, 4)
    l = R.get_model_config('large')
    assert (l.hidden_dim, l.num_layers, l.num_heads, l.ffn_dim) == (512, 12, 8, 2048)
    assert type(R.get_model_config('base')) is R.OneTransConfig          # D8 alias
    with pytest.raises(ValueError):
        R.get_model_config('nope')
    d = c.to_dict()
    c2 = R.OneTransConfig.from_dict({**d, 'hidden_dim': 128, 'unknown_key': 1})
    assert c2.hidden_dim == 128 and not hasattr(c2, 'unknown_key')
    assert len(c.ns_features) == 11


def test_scheduler_matches_reference_and_oracle():
    cfg = R.get_model_config('small')
    sch = R.PyramidScheduler(cfg)
    for L0 in [1, 2, 21, 272, 544, 1202, 2048]:
        for l in range(10):
            got = sch.get_layer_config(l, L0)
            want = O.reference_query_indices(l, L0, cfg.pyramid_ratios)
            assert got['query_indices'] == want
            if want is not None:
                assert got['keep_len'] == len(want)
    assert keep_lens_reference_ratio(544, 6, cfg.pyramid_ratios) == O.keep_lens_reference_ratio(544, 6, cfg.pyramid_ratios)
    assert keep_lens_linear_to_ns(544, 6, 32) == [458, 373, 288, 202, 117, 32]
    assert keep_lens_halving(2048, 6, 32) == [1024, 512, 256, 128, 64, 32]
    cfg.pyramid_enabled = False
    assert sch.get_layer_config(0, 100)['query_indices'] is None
    assert R.resolve_keep_lens(cfg, 100) == [100] * 6
    cfg.pyramid_enabled = True
    cfg.pyramid_keep_lens = [50, 60, 10, 10, 10, 3]
    assert R.resolve_keep_lens(cfg, 55) == [50, 50, 10, 10, 10, 3]     # clamped to the current length


@pytest.mark.parametrize('alignment', ['tail', 'head_literal'])
def test_position_segments_match_oracle_rule(alignment):
    """The segment table handed to the grouped GEMM covers every position exactly once with the group the
    per-position rule of OT/model.py:67-74 (or its D4 repair) assigns."""
    B = 3
    for L_ns in (1, 4, 16):
        for cur in (1, 3, 5, 16, 17, 40):
            for p0 in sorted({0, cur // 3, max(cur - L_ns, 0), max(cur - 2, 0), cur - 1}):
                segs = ops.position_segments(p0, cur, cur, L_ns, alignment, B)
                got = {}
                for (row_start, n_units, rpu, g0, gs) in segs:
                    for u in range(n_units):
                        for r in range(rpu):
                            row = row_start + u * rpu + r
                            assert row not in got
                            got[row] = g0 + u * gs
                assert sorted(got) == list(range((cur - p0) * B))
                for p in range(p0, cur):
                    for b in range(B):
                        assert got[(p - p0) * B + b] == O.group_of_position(p, cur, L_ns, alignment)


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert lib.ot_version() == _lib.ABI_VERSION == int(re.search(r'#define OT_ABI_VERSION (\d+)', open(HEADER).read()).group(1))
    declared = set(re.findall(r'^\s*(?:int|const char\*)\s+(ot_\w+)\s*\(', open(HEADER).read(), re.M))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_ctypes_structs_match_header_layout():
    """Compile a C probe that prints sizeof/offsetof for every params struct and compare with ctypes."""
    pairs = [('ot_gemm_seg', _lib.GemmSeg), ('ot_gemm_params', _lib.GemmParams), ('ot_ffn_params', _lib.FfnParams), ('ot_wgrad_seg', _lib.WgradSeg),
             ('ot_wgrad_params', _lib.WgradParams), ('ot_attn_params', _lib.AttnParams), ('ot_attn_cached_params', _lib.AttnCachedParams), ('ot_rmsnorm_params', _lib.RmsnormParams),
             ('ot_ns_tokenizer_params', _lib.NsTokenizerParams), ('ot_colsum_params', _lib.ColsumParams),
             ('ot_rmsprop_params', _lib.RmspropParams), ('ot_embed_params', _lib.EmbedParams), ('ot_heads_params', _lib.HeadsParams),
             ('ot_metrics_params', _lib.MetricsParams), ('ot_auc_params', _lib.AucParams)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for cname, st in pairs:
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname.rstrip("_")}));')
    lines += ['return 0;}']
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, 'p.c'), os.path.join(td, 'p')
        open(src, 'w').write('\n'.join(lines))
        subprocess.check_call(['gcc', '-std=c99', '-pedantic', '-Werror', '-o', exe, src])      # the header is plain C
        out = dict(l.split() for l in subprocess.check_output([exe]).decode().splitlines())
    for cname, st in pairs:
        assert int(out[cname]) == ctypes.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(out[f'{cname}.{fname}']) == getattr(st, fname).offset, (cname, fname)


def test_modules_refuse_cpu_tensors():
    cfg = R.get_model_config('small')
    cfg.num_layers, cfg.num_ns_tokens = 1, 2
    m = R.OneTransModel(cfg)
    non_seq, seq, _ = O.synthetic_batch(O.small_config(num_ns_tokens=2), 2, (3, 3, 3))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(non_seq, seq)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        R.RMSNorm(8)(torch.zeros(2, 8))


def test_param_name_map_roundtrip_and_counts():
    ocfg = O.small_config(num_ns_tokens=32)
    cfg = R.get_model_config('small')
    cfg.num_ns_tokens = 32
    m = R.OneTransModel(cfg)
    info = m.get_model_info()
    assert info['total_parameters'] == O.count_params(O.init_params(O.small_config(num_ns_tokens=32))) 
    assert abs(info['total_parameters'] / 1e6 - 143.6) < 0.1
    P = O.init_params(ocfg, seed=3)
    R.load_reference_style_params(m, P)
    E = R.export_reference_style_params(m)
    assert set(E) == set(P)
    assert all(torch.equal(E[k], P[k].reshape(E[k].shape)) for k in P)


def test_gelu_approximation_constants():
    """The epilogue GELU evaluates erf(x/sqrt2) as tanh(x*(c1 + c3 x^2 + c5 x^4)) (ot_common.cuh); check the constants'
    error bounds against the exact erf form the reference uses (Keras activation='gelu')."""
    import math
    src = open(os.path.join(ROOT, 'recommend_b200', 'csrc', 'ot_common.cuh')).read()
    m = re.search(r'kGeluC1 = ([-0-9.e+]+)f, kGeluC3 = ([-0-9.e+]+)f, kGeluC5 = ([-0-9.e+]+)f', src)
    c1, c3, c5 = (float(v) for v in m.groups())
    x = torch.linspace(-12, 12, 240001, dtype=torch.float64)
    xc = x.clamp(-8, 8)
    t = torch.tanh(xc * (c1 + c3 * xc ** 2 + c5 * xc ** 4))
    erf = torch.erf(x / math.sqrt(2))
    assert (t - erf).abs().max() < 4e-5
    gelu = 0.5 * x * (1 + t)
    assert (gelu - 0.5 * x * (1 + erf)).abs().max() < 6e-5
    du = c1 + 3 * c3 * xc ** 2 + 5 * c5 * xc ** 4
    grad = 0.5 * (1 + t) + xc * 0.5 * (1 - t * t) * du
    ref = 0.5 * (1 + erf) + x * torch.exp(-x * x / 2) / math.sqrt(2 * math.pi)
    assert (grad - ref).abs().max() < 1.5e-4


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under recommend_b200/ may import it, and bench.py only in its CPU legs."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in glob.glob(os.path.join(root, 'recommend_b200', '**', '*.py'), recursive=True):
        src = open(path).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), path
    bench = open(os.path.join(root, 'bench.py')).read()
    imports = [m.start() for m in re.finditer(r'^\s*from oracle import', bench, re.M)]
    cpu_legs = ('def cpu_oracle_samples_per_sec', 'def c5_cpu_candidates_per_sec')     # the two cpu_baseline / --impl reference timers
    assert len(imports) == len(cpu_legs), 'bench.py may touch the oracle only inside its CPU-baseline timers'
    for pos in imports:
        enclosing = max((bench.rfind(d, 0, pos), d) for d in ['def '] )[0]
        last_def = bench.rfind('\ndef ', 0, pos)
        assert any(bench.startswith(leg, last_def + 1) for leg in cpu_legs), \
            'bench.py may touch the oracle only inside cpu_oracle_samples_per_sec / c5_cpu_candidates_per_sec (cpu_baseline, --impl reference)'


def test_sample_batch_generator_matches_the_oracles():
    """Both bench arms must see the same synthetic inputs: the product generator (recommend_b200.data, after
    OT/data_loader.py:301-329) and the oracle's draw the same stream."""
    from recommend_b200.data import create_sample_batch
    cfg = R.get_model_config('small')
    cfg.num_ns_tokens = 16
    for mode in ('normal', 'ids'):
        a = create_sample_batch(cfg, 5, (4, 3, 2), seed=77, ns_mode=mode)
        b = O.synthetic_batch(O.small_config(num_ns_tokens=16), 5, (4, 3, 2), seed=77, ns_mode=mode)
        for da, db in zip(a, b):
            assert da.keys() == db.keys()
            for k in da:
                assert torch.equal(da[k], db[k]), (mode, k)


def test_public_surface_matches_the_reference_module():
    """Names, call signatures and return shapes a user of OT/model.py relies on (OT/__init__.py:9-26,
    OT/model.py:280-302, 395-416)."""
    import inspect
    for name in ('OneTransModel', 'OneTransConfig', 'get_model_config', 'RMSNorm', 'MixedMHA', 'MixedFFN', 'OneTransBlock', 'Tokenizer',
                 'PyramidScheduler', 'create_onetrans_model'):
        assert hasattr(R, name), name
    cfg = R.get_model_config('small')
    sched = R.PyramidScheduler(cfg).get_layer_config(0, 272)
    assert set(sched) == {'keep_ratio', 'query_indices', 'keep_len'} and sched['keep_len'] == 136
    assert list(sched['query_indices']) == list(range(136, 272))
    cfg.num_layers, cfg.num_ns_tokens = 1, 2
    m = R.create_onetrans_model('small')
    assert isinstance(m, R.OneTransModel)
    sig = inspect.signature(R.OneTransModel.forward)
    assert list(sig.parameters)[:5] == ['self', 'non_seq_features', 'seq_features', 'training', 'use_kv_cache']
    small = R.OneTransModel(cfg)
    info = small.get_model_info()
    assert set(info) >= {'total_parameters', 'num_layers', 'hidden_dim', 'num_heads'}
    assert info['total_parameters'] == sum(p.numel() for p in small.parameters()) and info['num_layers'] == 1
    small.reset_kv_cache()
    assert small.kv_cache is None
    for meth in ('build_kv_cache', 'score_candidates', 'forward_with_loss'):
        assert callable(getattr(small, meth))


# ---- rank-4 widening: datasets, Keras weight order, checkpoint files (host logic, CPU) ----------------------------------
def test_dataset_and_loader_surface():
    import recommend_b200 as R
    cfg = R.get_model_config('small')
    cfg.max_seq_len, cfg.batch_size = 8, 16
    dl = R.DataLoader(cfg)
    with pytest.raises(ValueError):
        dl.get_train_dataset()                                    # OT/data_loader.py:246-247
    dl.load_datasets(num_samples=(40, 20, 10))
    assert dl.get_data_info() == {'train_samples': 40, 'val_samples': 20, 'test_samples': 10}
    train = dl.get_train_dataset()
    assert len(train) == 3
    b1, b2 = [next(iter(train)) for _ in range(2)]               # re-iterable, reshuffled every epoch
    assert b1[0]['price'].shape == (16, 1) and b1[1]['click_seq'].shape == (16, 8, 64) and b1[2]['ctr'].shape == (16, 1)
    assert not torch.equal(b1[0]['price'], b2[0]['price'])
    t1, t2 = [next(iter(dl.get_test_dataset(4))) for _ in range(2)]
    assert torch.equal(t1[0]['price'], t2[0]['price']) and t1[0]['price'].shape == (4, 1)
    assert sum(b[2]['ctr'].shape[0] for b in dl.get_val_dataset(16)) == 20          # ragged last batch kept
    ns, seq, lab = dl.train_dataset[5]
    n = int(dl.train_dataset.seq_lens['cart_seq'][5])
    assert seq['cart_seq'].shape == (8, 64) and (seq['cart_seq'][:8 - n] == 0).all() and (seq['cart_seq'][8 - n:].abs().sum(1) > 0).all()
    sp = R.SequenceProcessor(cfg)                                                   # OT/data_loader.py:68-101
    assert sp.process_sequence(torch.zeros(0, 64)).shape == (8, 64)
    long = torch.arange(20.0).reshape(20, 1).expand(20, 64)
    assert torch.equal(sp.process_sequence(long)[:, 0], torch.arange(12.0, 20.0))   # the most recent events
    short = sp.process_sequence(torch.ones(3, 64))
    assert short[:5].abs().sum() == 0 and short[5:].eq(1).all()                     # padded in front
    assert set(sp.process_multi_sequences({'a': long, 'b': short})) == {'a', 'b'}


def test_keras_weight_order_and_checkpoint_roundtrip(tmp_path):
    import recommend_b200 as R
    from recommend_b200 import state
    cfg = R.get_model_config('small')
    cfg.num_layers, cfg.num_ns_tokens = 2, 3
    torch.manual_seed(0)
    m = R.OneTransModel(cfg)
    wl = state.keras_weight_list(m)
    names = [n for n, _ in wl]
    # OT/model.py:308-333 / :206-222 / :169-184 / :29-57 / :128-147 attribute order
    assert names[:2] == ['tokenizer/ns_tokenizer/dense/kernel', 'tokenizer/ns_tokenizer/dense/bias']
    assert names[8] == 'tokenizer/sep_embedding/embeddings' and wl[8][1].shape == (1, 256)
    blk = [n for n in names if n.startswith('blocks/0/')]
    assert blk[:5] == ['blocks/0/norm1/scale', 'blocks/0/norm2/scale', 'blocks/0/attention/Wq_shared/kernel',
                       'blocks/0/attention/Wk_shared/kernel', 'blocks/0/attention/Wv_shared/kernel']
    assert blk[5:8] == [f'blocks/0/attention/Wq_dedicated/{j}/kernel' for j in range(3)] and blk[14] == 'blocks/0/attention/Wo/kernel'
    assert blk[15:19] == ['blocks/0/ffn/ffn_shared/dense/kernel', 'blocks/0/ffn/ffn_shared/dense/bias',
                          'blocks/0/ffn/ffn_shared/dense_1/kernel', 'blocks/0/ffn/ffn_shared/dense_1/bias']
    assert names[-4:] == ['task_heads/cvr/dense/kernel', 'task_heads/cvr/dense/bias', 'task_heads/cvr/dense_1/kernel', 'task_heads/cvr/dense_1/bias']
    assert dict(wl)['blocks/1/ffn/ffn_dedicated/2/dense/kernel'].shape == (256, 1024)            # Keras [in, out]
    assert sum(a.size for _, a in wl) == m.get_model_info()['total_parameters']                  # every parameter exactly once
    d = 256
    assert np.array_equal(dict(wl)['blocks/1/attention/Wk_dedicated/1/kernel'], m.blocks[1].attention.Wqkv[2, :, d:2 * d].detach().numpy())
    state.save_weights(m, tmp_path / 'w.npz')
    m2 = R.OneTransModel(cfg)
    with torch.no_grad():
        for p in m2.parameters():
            p.normal_()
    state.load_weights(m2, tmp_path / 'w.npz')
    assert all(torch.equal(a, b) for a, b in zip(m.parameters(), m2.parameters()))
    # the TensorFlow-side export: np.savez(path, *model.get_weights()) -> positional arr_i keys
    np.savez(tmp_path / 'tf_side.npz', *[a for _, a in wl])
    m3 = R.OneTransModel(cfg)
    with torch.no_grad():
        for p in m3.parameters():
            p.normal_()
    state.load_weights(m3, tmp_path / 'tf_side.npz')
    assert all(torch.equal(a, b) for a, b in zip(m.parameters(), m3.parameters()))
    with pytest.raises(ValueError):
        state.load_keras_weight_list(m2, [a for _, a in wl][:-1])
    cfg3 = R.get_model_config('small')
    cfg3.num_layers, cfg3.num_ns_tokens = 2, 4
    with pytest.raises(ValueError):
        state.load_weights(R.OneTransModel(cfg3), tmp_path / 'w.npz')


def test_evaluator_percentile_is_numpy_linear():
    from recommend_b200.evaluate import _percentile
    xs = [0.3, 0.1, 0.9, 0.5, 0.7, 0.2]
    for q in (0, 50, 95, 99, 100):
        assert _percentile(xs, q) == pytest.approx(float(np.percentile(xs, q)))


def test_unsupported_head_dims_are_errors_before_any_device_work():
    """Argument validation of the attention entry points returns before the first CUDA call, so it runs without a GPU: head dims
    other than 32 / 64 / 96 are an unsupported-shape error, never a fallback."""
    lib = _lib.load()
    p = _lib.AttnParams()
    for f in ('q', 'k', 'v', 'o', 'lse', 'd_o', 'dq', 'dk', 'dv', 'delta'):
        setattr(p, f, 4096)                                  # never dereferenced: the shape check comes first
    p.ldq = p.ldk = p.ldv = p.ldo = p.lddo = p.lddq = p.lddk = p.lddv = 128
    p.B, p.H, p.Lq, p.Lk = 2, 4, 4, 4
    for hd in (16, 48, 128):
        p.head_dim = hd
        for fn in (lib.ot_attn_fwd, lib.ot_attn_bwd):
            assert fn(ctypes.byref(p), None) == -2           # OT_ERR_UNSUPPORTED_SHAPE, never a fallback
            assert f'head_dim={hd}' in lib.ot_last_error_string().decode()


def test_create_sample_batch_accepts_the_reference_call_form():
    """OT/data_loader.py:301-329: ``create_sample_batch(batch_size=2, config=None)`` - ids, one random length per sequence."""
    cfg = R.get_model_config('small')
    cfg.max_seq_len = 9
    for args, kw in (((4, cfg), {}), ((), dict(batch_size=4, config=cfg))):
        non_seq, seq, labels = R.create_sample_batch(*args, **kw)
        assert set(non_seq) == set(cfg.ns_features) and all(v.shape == (4, 1) for v in non_seq.values())
        assert all(v.shape[0] == 4 and 1 <= v.shape[1] <= 9 and v.shape[2] == 64 for v in seq.values())
        assert float(non_seq['user_id'].max()) < 100 and float(non_seq['item_id'].max()) < 1000 and (non_seq['user_id'] == non_seq['user_id'].round()).all()
        assert set(labels) == set(cfg.tasks)
    a, b = R.create_sample_batch(cfg, 5, (3, 2, 1)), R.create_sample_batch(config=cfg, batch_size=5, seq_lens=(3, 2, 1))
    assert all(torch.equal(a[1][k], b[1][k]) for k in a[1]) and a[1]['cart_seq'].shape == (5, 2, 64)


def test_command_lines_keep_the_reference_flags():
    """OT/train.py:378-419 and OT/evaluate.py:419-465: same flags; without a CUDA device they stop with exit code 2 (no CPU fallback)."""
    from recommend_b200 import train as T, evaluate as E
    for cli, argv in ((T._cli, ['--config', 'small', '--epochs', '1', '--batch_size', '8', '--model_dir', '/tmp/x', '--data_dir', '/tmp/d']),
                      (E._cli, ['--model_path', '/tmp/x', '--data_dir', '/tmp/d', '--output_dir', '/tmp/o', '--eval_type', 'offline'])):
        if not torch.cuda.is_available():
            assert cli(argv) == 2
        with pytest.raises(SystemExit) as e:
            cli(['--no-such-flag'])
        assert e.value.code == 2
    with pytest.raises(SystemExit):
        E._cli([])                                   # --model_path is required


def test_polynomial_exp2_building_block():
    """``ex2_poly`` of ot_common.cuh (FMA-pipe 2^x for the attention softmax loops; not wired into a kernel yet): its constants and
    its exact fp32 instruction sequence, emulated bit by bit, stay within 8e-5 of exp2 on the range a softmax uses."""
    src = open(os.path.join(os.path.dirname(_lib.LIB_PATH), '..', 'csrc', 'ot_common.cuh')).read()
    C = [np.float32(re.search(rf'#define OT_EX2_POLY_C{i} ([0-9.]+)f', src).group(1)) for i in range(4)]
    assert '12582912.0f' in src and '<< 23' in src

    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)

    def ex2_poly(x):
        x = np.maximum(x.astype(np.float32), np.float32(-126.0))
        M = np.float32(12582912.0)
        t = (x + M).astype(np.float32)
        f = (x - (t - M).astype(np.float32)).astype(np.float32)
        p = fma(f, np.full_like(f, C[3]), np.full_like(f, C[2]))
        p = fma(p, f, np.full_like(f, C[1]))
        p = fma(p, f, np.full_like(f, C[0]))
        bits = (p.view(np.int32).astype(np.int64) + ((t.view(np.int32).astype(np.int64) << 23) & 0xFFFFFFFF)) & 0xFFFFFFFF
        return bits.astype(np.uint32).view(np.float32), f

    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-125, 0, 400000), rng.uniform(-20, 0, 400000), -np.arange(0, 126, 0.5)]).astype(np.float32)
    y, f = ex2_poly(x)
    ref = np.exp2(x.astype(np.float64))
    assert np.abs(f).max() <= 0.5
    assert (np.abs(y.astype(np.float64) - ref) / ref).max() < 8e-5
    assert ex2_poly(np.array([-1000.0], np.float32))[0][0] < 2e-38          # far below the running maximum: as good as zero


def test_driver_facing_scripts_compile():
    """bench.py, __graft_entry__.py, the profiling scripts and the examples must at least be valid Python (the GPU arm cannot run here)."""
    import glob, py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, 'bench.py'), os.path.join(root, '__graft_entry__.py')] + glob.glob(os.path.join(root, 'profiles', '*.py')) + \
        glob.glob(os.path.join(root, 'examples', '*.py')) + glob.glob(os.path.join(root, 'recommend_b200', '*.py'))
    assert len(files) > 20
    for f in files:
        py_compile.compile(f, doraise=True)
