"""CPU tests of the oracle (oracle/onetrans_oracle.py): the KATs of SURVEY.md §4.3 that pin it, since the
reference ships no tests or golden vectors of its own; tests/test_reference_golden.py adds vectors produced by executing
the reference's own model code over a TensorFlow-op shim."""
import json
import math
import os

import pytest
import torch

from oracle import onetrans_oracle as O

RATIOS = [0.5, 0.3, 0.2, 0.1, 0.05, 0.03, 0.02, 0.01]   # OT/config.py:30


# ---- T1: index selection, bit-exact ------------------------------------------------------------------

def test_t1_literal_vectors():
    """Literal integer vectors of SURVEY.md §4.3 T1 (OT/model.py:292-296, OT/config.py:30)."""
    assert O.keep_lens_reference_ratio(272, 6, RATIOS) == [136, 81, 54, 27, 13, 8]
    assert O.keep_lens_reference_ratio(544, 6, RATIOS) == [272, 163, 108, 54, 27, 16]
    assert O.keep_lens_reference_ratio(1202, 8, RATIOS) == [601, 360, 240, 120, 60, 36, 24, 12]
    assert O.keep_lens_linear_to_ns(544, 6, 32) == [458, 373, 288, 202, 117, 32]
    assert O.keep_lens_linear_to_ns(272, 6, 16) == [229, 186, 144, 101, 58, 16]
    assert O.keep_lens_linear_to_ns(544, 8, 32) == [480, 416, 352, 288, 224, 160, 96, 32]
    assert O.keep_lens_halving(2048, 6, 32) == [1024, 512, 256, 128, 64, 32]
    assert O.keep_lens_halving(2048, 8, 32) == [1024, 512, 256, 128, 64, 32, 32, 32]


def test_t1_sweep_matches_reference_expression():
    """keep_len = max(1, int(total*ratio)) evaluated in Python double; query_indices = the tail."""
    for L0 in list(range(1, 300)) + [544, 1202, 2048, 4096]:
        for l, r in enumerate(RATIOS):
            k = O.reference_keep_len(l, L0, RATIOS)
            assert k == max(1, int(L0 * r))
            assert O.reference_query_indices(l, L0, RATIOS) == list(range(L0 - k, L0))
        assert O.reference_keep_len(len(RATIOS), L0, RATIOS) is None
    # smoke shapes of the reference's __main__ (OT/model.py:420-442): L = 10+1+5+1+4 = 21 -> [10, 6]
    assert O.keep_lens_reference_ratio(21, 2, RATIOS) == [10, 6]


def test_reference_model_smoke_main_shapes():
    """B=2, 2 layers, 4 NS tokens, click 10x64, cart 5x64 -> L_S=17, L=21."""
    cfg = O.OracleConfig(hidden_dim=128, num_layers=2, num_ns_tokens=4, ffn_dim=256,
                         ns_feature_names=['user_id', 'item_id', 'price'])
    P = O.init_params(cfg)
    g = torch.Generator().manual_seed(0)
    non_seq = {'user_id': torch.randint(0, 100, (2, 1), generator=g), 'item_id': torch.randint(0, 1000, (2, 1), generator=g),
               'price': torch.rand(2, 1, generator=g)}
    seq = {'click_seq': torch.rand(2, 10, 64, generator=g), 'cart_seq': torch.rand(2, 5, 64, generator=g)}
    x = O.tokenizer_forward(P, cfg, non_seq, seq)
    assert x.shape == (2, 21, 128)
    out = O.model_forward(P, cfg, non_seq, seq)
    assert set(out) == {'ctr', 'cvr'} and out['ctr'].shape == (2, 1)
    assert all(((v > 0) & (v < 1)).all() for v in out.values())


# ---- T2: causal mask ----------------------------------------------------------------------------------

def test_t2_mask_equals_band_part():
    """allowed(q,k) <=> k <= q + (Lk-Lq); equals tril(ones) when Lq == Lk (tf.linalg.band_part(ones,-1,0))."""
    cfg = O.OracleConfig(hidden_dim=16, num_heads=2, num_layers=1, num_ns_tokens=2, ffn_dim=32)
    P = O.init_params(cfg, dtype=torch.float64)
    torch.manual_seed(0)
    x = torch.randn(2, 7, 16, dtype=torch.float64)
    full = O.mixed_mha(P, 'blocks.0.attention.', cfg, x, 7)
    # recompute with an explicit tril mask, literal reference formulation
    pref = 'blocks.0.attention.'
    q = O._mixed_linear(x, P[pref + 'Wq'], None, 0, 7, 2, 'tail', False).reshape(2, 7, 2, 8)
    k = O._mixed_linear(x, P[pref + 'Wk'], None, 0, 7, 2, 'tail', False).reshape(2, 7, 2, 8)
    v = O._mixed_linear(x, P[pref + 'Wv'], None, 0, 7, 2, 'tail', False).reshape(2, 7, 2, 8)
    s = torch.einsum('bqhd,bkhd->bhqk', q, k) / math.sqrt(8.0)
    mask = torch.tril(torch.ones(7, 7))
    s = torch.where(mask == 1, s, torch.tensor(-1e9, dtype=torch.float64))
    o = torch.einsum('bhqk,bkhd->bqhd', torch.softmax(s, -1), v).reshape(2, 7, 16) @ P[pref + 'Wo']
    assert torch.allclose(full, o, atol=1e-12)
    tail = O.mixed_mha(P, pref, cfg, x, 3)
    assert torch.allclose(tail, full[:, 4:], atol=1e-12)


# ---- T3: tokenizer layout -----------------------------------------------------------------------------

def test_t3_tokenizer_layout_bit_exact():
    cfg = O.small_config(num_ns_tokens=4)
    P = O.init_params(cfg)
    O.randomize_small_params(P)
    non_seq, seq, _ = O.synthetic_batch(cfg, 3, (5, 4, 6))
    x = O.tokenizer_forward(P, cfg, non_seq, seq)
    assert x.shape == (3, 5 + 1 + 4 + 1 + 6 + 4, 256)
    sep = P['tokenizer.sep_embedding'][0]
    assert torch.equal(x[:, 5], sep.expand(3, 256)) and torch.equal(x[:, 10], sep.expand(3, 256))
    click = seq['click_seq'] @ P['tokenizer.seq_projections.0.kernel'] + P['tokenizer.seq_projections.0.bias']
    assert torch.equal(x[:, :5], click)
    pur = seq['purchase_seq'] @ P['tokenizer.seq_projections.2.kernel'] + P['tokenizer.seq_projections.2.bias']
    assert torch.equal(x[:, 11:17], pur)
    feats = torch.cat([non_seq[n] for n in cfg.ns_features], -1)
    ns = (feats @ P['tokenizer.ns_tokenizer.kernel'] + P['tokenizer.ns_tokenizer.bias']).reshape(3, 4, 256)
    assert torch.equal(x[:, 17:], ns)
    # absent features (OT/model.py:249-251, 269-275): the last configured sequence missing -> trailing SEP (D15)
    x2 = O.tokenizer_forward(P, cfg, non_seq, {k: v for k, v in seq.items() if k != 'purchase_seq'})
    assert x2.shape[1] == 5 + 1 + 4 + 1 + 4 and torch.equal(x2[:, 10], sep.expand(3, 256))
    x3 = O.tokenizer_forward(P, cfg, {}, seq)
    assert torch.equal(x3[:, 17:], torch.zeros(3, 4, 256))


# ---- T4 / T5: algebraic identities ----------------------------------------------------------------------

@pytest.mark.parametrize('alignment', ['tail', 'head_literal'])
def test_t4_tail_only_equals_compute_all_then_gather(alignment):
    cfg = O.OracleConfig(hidden_dim=64, num_layers=4, num_heads=4, ffn_dim=128, num_ns_tokens=3, ns_param_alignment=alignment)
    P = O.init_params(cfg, dtype=torch.float64)
    O.randomize_small_params(P)
    non_seq, seq, _ = O.synthetic_batch(cfg, 3, (9, 7, 8), dtype=torch.float64)
    a = O.model_forward(P, cfg, non_seq, seq, return_logits=True)
    b = O.model_forward(P, cfg, non_seq, seq, return_logits=True, query_mode='literal_gather')
    for t in a:
        assert torch.allclose(a[t], b[t], atol=1e-10)


@pytest.mark.parametrize('alignment', ['tail', 'head_literal'])
def test_t5_grouped_equals_per_token_loop(alignment):
    cfg = O.OracleConfig(hidden_dim=64, num_layers=3, num_heads=4, ffn_dim=128, num_ns_tokens=3, ns_param_alignment=alignment)
    P = O.init_params(cfg, dtype=torch.float64)
    O.randomize_small_params(P)
    non_seq, seq, _ = O.synthetic_batch(cfg, 2, (6, 5, 4), dtype=torch.float64)
    a = O.model_forward(P, cfg, non_seq, seq, return_logits=True)
    b = O.model_forward(P, cfg, non_seq, seq, return_logits=True, literal_loop=True, query_mode='literal_gather')
    for t in a:
        assert torch.allclose(a[t], b[t], atol=1e-10)


def test_group_of_position():
    # tail alignment: the last L_ns positions of the ORIGINAL sequence keep their own weights
    assert [O.group_of_position(p, 6, 2, 'tail') for p in range(6)] == [0, 0, 0, 0, 1, 2]
    assert [O.group_of_position(p, 1, 2, 'tail') for p in range(1)] == [2]            # pruned below L_ns: token 1 survives
    assert [O.group_of_position(p, 6, 2, 'head_literal') for p in range(6)] == [1, 2, 0, 0, 0, 0]


# ---- T6: causality -----------------------------------------------------------------------------------

def test_t6_causality_bit_identical():
    cfg = O.OracleConfig(hidden_dim=32, num_layers=1, num_heads=2, ffn_dim=64, num_ns_tokens=2, pyramid_enabled=False)
    P = O.init_params(cfg)
    torch.manual_seed(1)
    x = torch.randn(2, 12, 32)
    y0 = O.block_forward(P, 0, cfg, x, 12)
    x2 = x.clone()
    x2[:, 7] += 1.0
    y1 = O.block_forward(P, 0, cfg, x2, 12)
    assert torch.equal(y0[:, :7], y1[:, :7]) and not torch.equal(y0[:, 7:], y1[:, 7:])


# ---- T7: RMSNorm -------------------------------------------------------------------------------------

def test_t7_rmsnorm_hand_computed():
    x = torch.tensor([[1.0, 2.0, 3.0, 4.0], [0.0, 0.0, 0.0, 0.0]], dtype=torch.float64)
    g = torch.tensor([1.0, 0.5, 2.0, 1.0], dtype=torch.float64)
    y = O.rmsnorm(x, g, 1e-6)
    r = 1.0 / math.sqrt((1 + 4 + 9 + 16) / 4 + 1e-6)
    assert torch.allclose(y[0], torch.tensor([1 * r, 2 * r * 0.5, 3 * r * 2.0, 4 * r], dtype=torch.float64), atol=1e-15)
    assert torch.equal(y[1], torch.zeros(4, dtype=torch.float64))
    # folded-gain identity used by fused kernels: rms(x)*g @ W == r * (x @ diag(g) W)
    torch.manual_seed(0)
    X = torch.randn(5, 8, dtype=torch.float64)
    G = torch.rand(8, dtype=torch.float64) + 0.5
    W = torch.randn(8, 3, dtype=torch.float64)
    rr = torch.rsqrt(X.square().mean(-1, keepdim=True) + 1e-6)
    assert torch.allclose(O.rmsnorm(X, G) @ W, rr * (X @ (G[:, None] * W)), atol=1e-12)


# ---- T11: BCE ----------------------------------------------------------------------------------------

def test_t11_bce_matches_keras_formula_and_logits_form():
    p = torch.tensor([[0.9], [0.2], [1.0], [0.0]], dtype=torch.float64)
    y = torch.tensor([[1.0], [0.0], [1.0], [1.0]], dtype=torch.float64)
    eps = 1e-7
    want = 0.0
    for pi, yi in zip(p.flatten().tolist(), y.flatten().tolist()):
        pc = min(max(pi, eps), 1 - eps)
        want += -(yi * math.log(pc + eps) + (1 - yi) * math.log(1 - pc + eps))
    want /= 4
    got = O.bce_loss({'ctr': p}, {'ctr': y}, ['ctr'])
    assert abs(float(got) - want) < 1e-12
    # away from the clip the probability form equals the logits form
    z = torch.tensor([[0.3], [-1.2], [2.0]], dtype=torch.float64)
    yy = torch.tensor([[1.0], [0.0], [0.0]], dtype=torch.float64)
    a = O.bce_loss({'ctr': torch.sigmoid(z)}, {'ctr': yy}, ['ctr'])
    b = torch.nn.functional.binary_cross_entropy_with_logits(z, yy)
    assert abs(float(a) - float(b)) < 1e-6


def test_clip_by_norm():
    g = torch.tensor([3.0, 4.0])
    assert torch.allclose(O.clip_by_norm(g, 10.0), g)
    assert torch.allclose(O.clip_by_norm(g, 2.5), g * 0.5)


def test_param_counts():
    """S with L_NS=32 has 143.6 M parameters, 23.90 M per block."""
    cfg = O.small_config(num_ns_tokens=32)
    d, F, G = 256, 1024, 33
    per_block = 2 * d + 3 * G * d * d + d * d + G * (d * F + F + F * d + d)
    assert abs(per_block / 1e6 - 23.90) < 0.01
    total = 6 * per_block + 11 * d * 32 + d * 32 + 3 * (64 * d + d) + d + d + 2 * (d * (d // 2) + d // 2 + d // 2 + 1)
    assert abs(total / 1e6 - 143.6) < 0.1


# ---- golden vectors generated by the oracle itself (tests/golden/make_golden.py) -----------------------

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'oracle_golden.json')


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason='golden file not generated')
def test_golden_vectors():
    from tests.golden.make_golden import cases, run_case
    want = json.load(open(GOLDEN))
    for name, spec in cases().items():
        got = run_case(spec)
        for k, v in want[name].items():
            assert torch.allclose(torch.tensor(got[k]), torch.tensor(v), rtol=1e-5, atol=1e-6), (name, k)


def test_T12_clip_and_rmsprop_hand_computed():
    """OT/train.py:133-138 with Keras-2.12 RMSprop arithmetic, two steps by hand."""
    g = torch.tensor([3.0, 4.0], dtype=torch.float64)
    assert torch.equal(O.clip_by_norm(g, 10.0), g)                       # ||g|| = 5 <= 10: untouched
    assert torch.allclose(O.clip_by_norm(g, 2.5), g * 0.5, atol=0, rtol=1e-15)
    w, rms, mom = torch.tensor([1.0, -1.0], dtype=torch.float64), torch.zeros(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
    lr, rho, m, eps = 0.005, 0.9, 0.5, 1e-7
    w1, rms1, mom1 = O.rmsprop_step(w, g, rms, mom, lr, rho, m, eps)
    r = [0.1 * 9.0, 0.1 * 16.0]
    inc = [lr * 3.0 / (r[0] + eps) ** 0.5, lr * 4.0 / (r[1] + eps) ** 0.5]
    assert torch.allclose(rms1, torch.tensor(r, dtype=torch.float64), rtol=1e-14)
    assert torch.allclose(mom1, torch.tensor(inc, dtype=torch.float64), rtol=1e-14)
    assert torch.allclose(w1, torch.tensor([1.0 - inc[0], -1.0 - inc[1]], dtype=torch.float64), rtol=1e-14)
    w2, rms2, mom2 = O.rmsprop_step(w1, g, rms1, mom1, lr, rho, m, eps)
    r2 = [0.9 * r[0] + 0.1 * 9.0, 0.9 * r[1] + 0.1 * 16.0]
    inc2 = [lr * 3.0 / (r2[0] + eps) ** 0.5, lr * 4.0 / (r2[1] + eps) ** 0.5]
    assert torch.allclose(mom2, torch.tensor([m * inc[0] + inc2[0], m * inc[1] + inc2[1]], dtype=torch.float64), rtol=1e-14)
    assert torch.allclose(w2, w1 - mom2, rtol=1e-14)
    # momentum 0: plain step, mom untouched
    w3, _, mom3 = O.rmsprop_step(w, g, rms, None, lr, rho, 0.0, eps)
    assert mom3 is None and torch.allclose(w3, w1, rtol=1e-14)


def test_T13_event_embedding_lookup_concat_and_adagrad():
    """ID front end (extension): lookup-and-concat is a copy; sparse gradient == summed IndexedSlices;
    Keras Adagrad by hand."""
    table = torch.arange(5 * 2, dtype=torch.float32).reshape(5, 2)          # field 0: rows 0-2, field 1: rows 3-4
    ids = torch.tensor([[[2, 1], [0, 0]], [[2, 0], [1, 1]]])                # [B=2, L=2, 2 fields]
    ev = O.embed_events(table, [3, 2], ids, out_dtype=torch.float32)
    assert ev.shape == (2, 2, 4)
    assert torch.equal(ev[0, 0], torch.tensor([4., 5., 8., 9.]))            # row 2 | row 3+1
    assert torch.equal(ev[1, 1], torch.tensor([2., 3., 8., 9.]))
    d = torch.ones(2, 2, 4)
    g = O.embed_grad_table(table, [3, 2], ids, d)
    assert torch.equal(g[:, 0], torch.tensor([1., 1., 2., 2., 2.], dtype=torch.float64))   # row 2 twice, rows 3 and 4 twice each
    w, acc = O.adagrad_step(torch.tensor([1.0]), torch.tensor([2.0]), torch.tensor([0.1]), lr=0.1, eps=1e-7)
    assert abs(float(acc) - 4.1) < 1e-6 and abs(float(w) - (1.0 - 0.1 * 2.0 / (4.1 ** 0.5 + 1e-7))) < 1e-6


# ---- two-stage inference restatement (oracle-internal pins, fp64) --------------------------------------------------------
def _two_stage_setup(schedule, pyramid=True, layers=3, L_ns=4):
    cfg = O.OracleConfig(hidden_dim=64, num_layers=layers, num_heads=4, ffn_dim=128, num_ns_tokens=L_ns, pyramid_enabled=pyramid)
    L0 = 9 + 7 + 11 + 2 + L_ns
    cfg.pyramid_keep_lens = {'linear_to_ns': O.keep_lens_linear_to_ns(L0, layers, L_ns), 'halving': O.keep_lens_halving(L0, layers, L_ns),
                             'reference_ratio': None}[schedule]
    P = O.init_params(cfg, seed=3, dtype=torch.float64)
    O.randomize_small_params(P, seed=4)
    non_seq, seq, _ = O.synthetic_batch(cfg, 5, (9, 7, 11), seed=8)
    to64 = lambda d: {k: v.double() for k, v in d.items()}
    return cfg, P, to64(non_seq), to64({k: v[:1] for k, v in seq.items()})


@pytest.mark.parametrize('schedule', ['linear_to_ns', 'reference_ratio', 'halving'])
def test_two_stage_equals_full_forward(schedule):
    cfg, P, non_seq, seq1 = _two_stage_setup(schedule)
    full = O.model_forward(P, cfg, non_seq, {k: v.expand(5, -1, -1) for k, v in seq1.items()}, return_logits=True)
    cache = O.two_stage_user_cache(P, cfg, seq1)
    got = O.two_stage_score(P, cfg, cache, non_seq)
    for t in cfg.tasks:
        assert torch.allclose(got[t], full[t], rtol=0, atol=1e-10)


def test_two_stage_extend_without_pyramid_is_a_fresh_build():
    cfg, P, non_seq, seq1 = _two_stage_setup('linear_to_ns', pyramid=False)
    last = cfg.sequence_features[-1]
    head = dict(seq1)
    head[last] = seq1[last][:, :6]
    cache = O.two_stage_extend(P, cfg, O.two_stage_user_cache(P, cfg, head), seq1[last][:, 6:])
    fresh = O.two_stage_user_cache(P, cfg, seq1)
    for a, b in zip(cache['layers'], fresh['layers']):
        assert torch.allclose(a[0], b[0], rtol=0, atol=1e-10) and torch.allclose(a[1], b[1], rtol=0, atol=1e-10)
    # the NS-side plan of a cache is the one it was built with: score both with the fresh plan
    got = O.two_stage_score(P, cfg, {**cache, 'plan': fresh['plan']}, non_seq)
    want = O.two_stage_score(P, cfg, fresh, non_seq)
    for t in cfg.tasks:
        assert torch.allclose(got[t], want[t], rtol=0, atol=1e-10)


def test_two_stage_extend_with_pyramid_only_grows_key_sets():
    cfg, P, non_seq, seq1 = _two_stage_setup('linear_to_ns')
    base = O.two_stage_user_cache(P, cfg, seq1)
    ext = O.two_stage_extend(P, cfg, base, torch.randn(1, 3, 64, dtype=torch.float64, generator=torch.Generator().manual_seed(2)))
    n = 3
    for (cur, Tn, cur_S, keep, Tq, keep_S), a, b in zip(base['plan'], base['layers'], ext['layers']):
        if a is None or n == 0:
            assert (a is None and b is None) or a[0].shape == b[0].shape
            continue
        assert b[0].shape[1] == a[0].shape[1] + n and torch.equal(b[0][:, :a[0].shape[1]], a[0])    # old keys untouched
        n = min(n, keep_S)
    assert not torch.allclose(O.two_stage_score(P, cfg, ext, non_seq)['ctr'], O.two_stage_score(P, cfg, base, non_seq)['ctr'])


def test_clip_is_per_keras_variable():
    """OT/train.py:133-135 clips per Keras variable: every weight-group index of the packed per-position Dense tensors is one."""
    g = torch.zeros(3, 2, 2, dtype=torch.float64)
    g[0] = 3.0          # ||.|| = 6
    g[1] = 0.5          # ||.|| = 1
    g[2, 0, 0] = 10.0   # ||.|| = 10
    out = O.clip_per_keras_variable('blocks.0.ffn.W1', g, 2.0)
    assert torch.allclose(out[0], g[0] * (2.0 / 6.0)) and torch.equal(out[1], g[1]) and torch.allclose(out[2], g[2] * 0.2)
    whole = O.clip_per_keras_variable('blocks.0.attention.Wo', g, 2.0)     # not packed: one variable
    assert torch.allclose(whole, g * (2.0 / float(g.norm())))
    P = {'blocks.0.ffn.b1': torch.zeros(2, 4, dtype=torch.float64), 'output_norm.scale': torch.ones(4, dtype=torch.float64)}
    G = {'blocks.0.ffn.b1': torch.tensor([[3.0, 4.0, 0, 0], [0.3, 0.4, 0, 0]], dtype=torch.float64), 'output_norm.scale': torch.full((4,), 5.0, dtype=torch.float64)}
    st = {}
    O.clip_rmsprop_update(P, G, st, lr=1.0, rho=0.0, momentum=0.0, eps=1e-7, clip_norm=1.0)     # rho = 0: rms == clipped g^2
    assert torch.allclose(st['blocks.0.ffn.b1']['rms'][0], torch.tensor([0.36, 0.64, 0, 0], dtype=torch.float64))   # (3,4)/5 clipped to norm 1
    assert torch.allclose(st['blocks.0.ffn.b1']['rms'][1], torch.tensor([0.09, 0.16, 0, 0], dtype=torch.float64))   # below the clip: untouched
