"""GPU parity of the model-level path (tokenizer -> blocks -> heads, forward and backward) against the CPU
oracle (oracle/onetrans_oracle.py) on identical weights and inputs.  Tolerances are the north_star's bf16
tolerances: logits rel err <= 1e-2; gradients are compared by relative L2 error per tensor."""
import pytest
import torch

from oracle import onetrans_oracle as O
import recommend_b200 as R
from tests.helpers import make_configs, oracle_keep_lens, bf16_round_inputs, to_cuda, rel_err, rel_l2

pytestmark = pytest.mark.gpu

# north_star: logits rel err <= 1e-2 (bf16 kernels vs fp32 oracle); per-tensor relative L2 error of parameter gradients <= 3e-2
# (bf16 activations / gradients).  The environment can only TIGHTEN them (experiments with a stricter bar), never loosen them.
LOGIT_TOL = min(1e-2, float(__import__('os').environ.get('OT_LOGIT_TOL', 1e-2)))
GRAD_L2_TOL = min(3e-2, float(__import__('os').environ.get('OT_GRAD_TOL', 3e-2)))
# element-wise companion of the aggregate bar: worst single logit error relative to the largest logit.  The aggregate is an
# average over the batch, the worst of n logits sits ~sqrt(2 ln n) standard deviations out (2.9x at n = 64, 3.7x at n = 4096),
# so it is held to 2.5x the aggregate bar; profiles/exp_rounding_ablation.py shows where the bf16 floor (0.7e-2 aggregate) comes from.
LOGIT_MAX_TOL = 2.5e-2


def _build(ocfg, cfg, seed=0):
    P = O.init_params(ocfg, seed=seed)
    O.randomize_small_params(P, seed=seed + 1)
    # "identical inputs and weights": the GEMM weights are made bf16-representable on BOTH sides (as a bf16
    # checkpoint would be), so the comparison measures the kernels' arithmetic, not the one-off weight cast.
    for k in P:
        if P[k].dim() >= 2 and 'ns_tokenizer' not in k and 'task_heads' not in k and 'sep_embedding' not in k:
            P[k] = P[k].to(torch.bfloat16).to(torch.float32)
    model = R.OneTransModel(cfg).cuda()
    R.load_reference_style_params(model, P)
    return P, model


def _run_pair(ocfg, cfg, B, seq_lens, seed=0, present=None, with_grads=True):
    P, model = _build(ocfg, cfg, seed)
    non_seq, seq, labels = O.synthetic_batch(ocfg, B, seq_lens, seed=1234 + seed)
    if present is not None:
        seq = {k: v for k, v in seq.items() if k in present}
    non_seq, seq = bf16_round_inputs(non_seq, seq)
    L0 = sum(v.shape[1] for v in seq.values()) + ocfg.num_ns_tokens
    n_seq = len(ocfg.sequence_features)
    L0 += sum(1 for i, n in enumerate(ocfg.sequence_features) if n in seq and i < n_seq - 1)
    oracle_keep_lens(ocfg, cfg, L0)
    if with_grads:
        loss_o, grads_o, _ = O.loss_and_grads(P, ocfg, non_seq, seq, labels)
    logits_o = O.model_forward(P, ocfg, non_seq, seq, return_logits=True)
    model.zero_grad(set_to_none=True)
    logits_g = model(to_cuda(non_seq), to_cuda(seq), training=with_grads, return_logits=True)
    out = {'logits_o': logits_o, 'logits_g': {k: v.detach().float().cpu() for k, v in logits_g.items()}}
    if with_grads:
        probs = {k: torch.sigmoid(v) for k, v in logits_g.items()}
        eps = 1e-7
        loss = 0.0
        for t in cfg.tasks:   # Keras BinaryCrossentropy(from_logits=False) (OT/train.py:84-87)
            p = probs[t].clamp(eps, 1 - eps)
            y = labels[t].cuda()
            loss = loss + (-(y * torch.log(p + eps) + (1 - y) * torch.log(1 - p + eps))).mean()
        loss.backward()
        torch.cuda.synchronize()
        out.update(loss_o=float(loss_o), loss_g=float(loss.detach()), grads_o=grads_o,
                   grads_g=R.export_reference_style_params(model, grads=True))
    return out


def _check(out, logit_tol=LOGIT_TOL, grad_tol=GRAD_L2_TOL):
    lo = torch.cat([out['logits_o'][t].flatten() for t in out['logits_o']])
    lg = torch.cat([out['logits_g'][t].flatten() for t in out['logits_o']])
    e = rel_l2(lg, lo)   # "logits rel err": ||gpu - oracle||_2 / ||oracle||_2 over the batch and tasks
    em = rel_err(lg, lo)
    print(f'logits rel-L2 err {e:.3e} (max-abs/max {em:.3e})')
    assert e <= logit_tol, f'logits rel err {e:.3e}'
    assert em <= LOGIT_MAX_TOL, f'worst logit error / largest logit {em:.3e}'
    if 'grads_o' in out:
        assert abs(out['loss_g'] - out['loss_o']) <= 1e-2 * max(1.0, abs(out['loss_o']))
        worst = {}
        for k, go in out['grads_o'].items():
            gg = out['grads_g'][k]
            if go.abs().max() == 0:      # parameter not on the path (absent sequence): no or zero gradient
                assert gg is None or gg.abs().max() < 1e-6, k
                continue
            assert gg is not None, f'no gradient for {k}' 
            worst[k] = rel_l2(gg.reshape(go.shape), go)
        print('worst gradient rel-L2:', sorted(worst.items(), key=lambda kv: -kv[1])[:4])
        bad = {k: v for k, v in worst.items() if not v <= grad_tol}
        assert not bad, f'gradient rel-L2 errors above {grad_tol}: {bad}'


def test_smoke_shape_of_reference_main():
    """The reference's own __main__ smoke shapes (OT/model.py:420-442): d=128 -> here d=256 H=4 (head_dim 64),
    2 layers, 4 NS tokens, click 10 + cart 5 events.  Batch 64 instead of the script's 2: with four logits the relative
    L2 error is a coin toss around the bf16 floor (0.8-1.1e-2 depending on the rounding of a single row statistic)."""
    ocfg, cfg = make_configs(hidden_dim=256, num_layers=2, ffn_dim=512, num_ns_tokens=4)
    _check(_run_pair(ocfg, cfg, B=64, seq_lens=(10, 5, 7), present=('click_seq', 'cart_seq')))


@pytest.mark.parametrize('case', ['G_d256_pyramid_on_1_block', 'H_d256_pyramid_off_2_blocks'])
def test_product_equals_the_reference_outputs(case):
    """DIRECT product-vs-reference parity: the expected probabilities were produced by the reference's own, unmodified
    ``OneTransModel.call`` (OT/model.py:335-393) executed over oracle/tf_shim.py on these weights and inputs
    (tests/golden/make_reference_golden.py; the CPU twin in tests/test_reference_golden.py shows that the seeds rebuild what the
    reference saw).  The reference gives dedicated weights to positions < num_ns_tokens: ``ns_param_alignment='head_literal'``."""
    import json, os
    import numpy as np
    from tests.helpers import REFERENCE_KERNEL_CASES, reference_case_inputs, reference_case_checksum
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    Z = np.load(os.path.join(here, 'reference_golden.npz'))
    spec = json.load(open(os.path.join(here, 'reference_golden.json')))['cases'][case]
    assert {k: (tuple(v) if isinstance(v, list) else v) for k, v in spec.items() if k in REFERENCE_KERNEL_CASES[case]} == REFERENCE_KERNEL_CASES[case]
    ocfg, P, non_seq, seq = reference_case_inputs(spec)
    assert reference_case_checksum(P, non_seq, seq) == pytest.approx(spec['checksum'], rel=1e-13)
    _, cfg = make_configs(hidden_dim=spec['hidden_dim'], num_layers=spec['num_layers'], num_heads=spec['num_heads'], ffn_dim=spec['ffn_dim'],
                          num_ns_tokens=spec['num_ns_tokens'], schedule='reference_ratio', alignment='head_literal',
                          pyramid_enabled=spec['pyramid_enabled'])
    model = R.OneTransModel(cfg).cuda()
    R.load_reference_style_params(model, P)
    with torch.no_grad():
        got = model(to_cuda(non_seq), to_cuda(seq), training=False, return_logits=True)
    ref_prob = torch.cat([torch.from_numpy(Z[f'{case}/out/prob/{t}']).flatten() for t in cfg.tasks])
    lo = torch.log(ref_prob) - torch.log1p(-ref_prob)                     # the reference returns probabilities (sigmoid heads, OT/model.py:329)
    lg = torch.cat([got[t].flatten().double().cpu() for t in cfg.tasks])
    e = rel_l2(lg, lo)
    print(f'product vs REFERENCE [{case}] logits rel-L2 err {e:.3e}')
    assert e <= LOGIT_TOL, f'logits rel err {e:.3e}'
    assert (torch.sigmoid(lg) - ref_prob).abs().max() < 1e-2


def test_reference_example_shape_d128_head_dim_32():
    """The shape of the reference's own smoke block and example scripts (OT/model.py:420-442, OT/examples/train_example.py:22-27):
    hidden_dim 128 with 4 heads = head_dim 32, 2 layers, 4 NS tokens, click 10 + cart 5 events (batch 64 for a stable statistic)."""
    ocfg, cfg = make_configs(hidden_dim=128, num_layers=2, num_heads=4, ffn_dim=512, num_ns_tokens=4)
    _check(_run_pair(ocfg, cfg, B=64, seq_lens=(10, 5, 7), present=('click_seq', 'cart_seq')))


def test_c1_small_reference_ratio():
    """BASELINE config 1 shapes: OneTrans-S, B=32, 256 S + 16 NS tokens, reference ratio schedule."""
    ocfg, cfg = make_configs(num_ns_tokens=16, schedule='reference_ratio')
    _check(_run_pair(ocfg, cfg, B=32, seq_lens=(86, 84, 84)))


def test_c1_small_linear_to_ns():
    ocfg, cfg = make_configs(num_ns_tokens=16, schedule='linear_to_ns')
    _check(_run_pair(ocfg, cfg, B=32, seq_lens=(86, 84, 84)))


def test_head_literal_alignment():
    """ns_param_alignment='head_literal' = OT/model.py:69-74 as written (D4)."""
    ocfg, cfg = make_configs(num_layers=3, num_ns_tokens=8, alignment='head_literal', schedule='linear_to_ns')
    _check(_run_pair(ocfg, cfg, B=16, seq_lens=(40, 30, 20)))


def test_batch_multiple_of_128_and_ragged_batch():
    ocfg, cfg = make_configs(num_layers=2, num_ns_tokens=8, schedule='linear_to_ns')
    _check(_run_pair(ocfg, cfg, B=256, seq_lens=(30, 20, 10)))
    ocfg, cfg = make_configs(num_layers=2, num_ns_tokens=8, schedule='linear_to_ns')
    _check(_run_pair(ocfg, cfg, B=200, seq_lens=(30, 20, 10), seed=3))


def test_onetrans_l_shapes():
    """OneTrans-L (d=384, H=4 -> head_dim 96, F=1536), shortened to 3 blocks."""
    ocfg, cfg = make_configs(hidden_dim=384, num_layers=3, ffn_dim=1536, num_ns_tokens=8, schedule='linear_to_ns')
    _check(_run_pair(ocfg, cfg, B=24, seq_lens=(60, 50, 40)))


def test_c3_onetrans_l_full_depth():
    """BASELINE config 3 at full depth: OneTrans-L = ``get_model_config('default')`` (OT/config.py:14-17: d 384, 8 blocks, 4 heads ->
    head_dim 96, F 1536), 512 S + 32 NS tokens, linear_to_ns = [480, 416, 352, 288, 224, 160, 96, 32]; batch 128 instead of 2048
    per GPU so that the fp32 oracle (forward + gradients) finishes in about a minute on the host cores."""
    ocfg, cfg = make_configs(hidden_dim=384, num_layers=8, ffn_dim=1536, num_ns_tokens=32, schedule='linear_to_ns')
    out = _run_pair(ocfg, cfg, B=128, seq_lens=(170, 170, 170))
    assert R.resolve_keep_lens(cfg, 544) == [480, 416, 352, 288, 224, 160, 96, 32]
    _check(out)


def test_onetrans_large_config():
    """``OneTransLargeConfig`` (OT/config.py:95-103: d 512, 12 blocks, 8 heads, F 2048 - not a paper configuration):
    N = 512 rows do not fit one epilogue tile, so the RMSNorms run as stand-alone kernels here."""
    c = R.get_model_config('large')
    assert (c.hidden_dim, c.num_layers, c.num_heads, c.ffn_dim) == (512, 12, 8, 2048)
    ocfg, cfg = make_configs(hidden_dim=512, num_layers=12, num_heads=8, ffn_dim=2048, num_ns_tokens=16, schedule='linear_to_ns')
    _check(_run_pair(ocfg, cfg, B=32, seq_lens=(60, 50, 40)))


def test_c4_long_sequence_halving():
    """BASELINE config 4 shapes: three behaviour sequences concatenated to 2048 tokens (2016 S + 32 NS), query set
    halved per block (1024, 512, ... 32); batch cut to 8 so that the fp32 oracle finishes in seconds."""
    ocfg, cfg = make_configs(num_ns_tokens=32, schedule='halving')
    out = _run_pair(ocfg, cfg, B=8, seq_lens=(672, 671, 671))
    assert R.resolve_keep_lens(cfg, 2048) == [1024, 512, 256, 128, 64, 32]
    _check(out)


def test_c2_sequence_length_linear_to_ns():
    """BASELINE config 2 token counts (512 S + 32 NS, linear_to_ns = [458, 373, 288, 202, 117, 32]) at batch 32."""
    ocfg, cfg = make_configs(num_ns_tokens=32, schedule='linear_to_ns')
    out = _run_pair(ocfg, cfg, B=32, seq_lens=(170, 170, 170))
    assert R.resolve_keep_lens(cfg, 544) == [458, 373, 288, 202, 117, 32]
    _check(out)


def test_pyramid_disabled_and_missing_sequence():
    ocfg, cfg = make_configs(num_layers=2, num_ns_tokens=4, pyramid_enabled=False)
    _check(_run_pair(ocfg, cfg, B=8, seq_lens=(20, 10, 5), present=('click_seq', 'purchase_seq')))


def test_causality():
    """T6: perturbing behaviour event j leaves hidden states at positions < j unchanged (bit-identical: the
    kernels are deterministic per row and never read later positions)."""
    ocfg, cfg = make_configs(num_layers=1, num_ns_tokens=4, pyramid_enabled=False)
    P, model = _build(ocfg, cfg)
    non_seq, seq, _ = O.synthetic_batch(ocfg, 4, (30, 20, 10))
    blk = model.blocks[0]
    with torch.no_grad():
        x = model.tokenizer(to_cuda(non_seq), to_cuda(seq))
        y0, _ = blk(x)
        x2 = x.clone()
        j = 25
        x2[:, j, :] += 1.0
        y1, _ = blk(x2)
    assert torch.equal(y0[:, :j, :], y1[:, :j, :])
    assert not torch.equal(y0[:, j:, :], y1[:, j:, :])


def test_t9_auc_delta_on_64k_samples():
    """north_star: AUC delta <= 1e-4 (exact ROC-AUC, sklearn) between the bf16 kernels and the fp32 oracle on >= 64k
    synthetic samples.  Labels are drawn from the oracle's own probabilities so that the AUC is meaningful."""
    from sklearn.metrics import roc_auc_score
    ocfg, cfg = make_configs(num_layers=2, ffn_dim=512, num_ns_tokens=8, schedule='linear_to_ns')
    P, model = _build(ocfg, cfg, seed=9)
    N, chunk = 65536, 4096
    seq_lens = (12, 10, 8)
    ocfg.pyramid_keep_lens = R.resolve_keep_lens(cfg, sum(seq_lens) + 2 + 8)
    po, pg = {t: [] for t in cfg.tasks}, {t: [] for t in cfg.tasks}
    for i in range(N // chunk):
        non_seq, seq, _ = O.synthetic_batch(ocfg, chunk, seq_lens, seed=100 + i)
        non_seq, seq = bf16_round_inputs(non_seq, seq)
        o = O.model_forward(P, ocfg, non_seq, seq)
        with torch.no_grad():
            g = model(to_cuda(non_seq), to_cuda(seq))
        for t in cfg.tasks:
            po[t].append(o[t].flatten())
            pg[t].append(g[t].flatten().float().cpu())
    gen = torch.Generator().manual_seed(0)
    for t in cfg.tasks:
        a, b = torch.cat(po[t]), torch.cat(pg[t])
        logit = torch.logit(a.clamp(1e-6, 1 - 1e-6)) * 3.0          # sharpen: random-init logits are small
        y = (torch.rand(N, generator=gen) < torch.sigmoid(logit)).numpy()
        auc_o, auc_g = roc_auc_score(y, a.numpy()), roc_auc_score(y, b.numpy())
        print(f'AUC[{t}] oracle {auc_o:.6f} kernels {auc_g:.6f} delta {abs(auc_o - auc_g):.2e}')
        assert abs(auc_o - auc_g) <= 1e-4


def test_t9_auc_delta_on_c1_shapes():
    """T9 on a BASELINE configuration: OneTrans-S at config 1's token counts (256 S + 16 NS, 6 blocks, reference ratio schedule),
    65536 samples in chunks of 2048; AUC delta <= 1e-4 (exact ROC-AUC)."""
    from sklearn.metrics import roc_auc_score
    ocfg, cfg = make_configs(num_ns_tokens=16, schedule='reference_ratio')
    P, model = _build(ocfg, cfg, seed=4)
    N, chunk = 65536, 2048
    seq_lens = (86, 84, 84)
    ocfg.pyramid_keep_lens = R.resolve_keep_lens(cfg, 272)
    po, pg = {t: [] for t in cfg.tasks}, {t: [] for t in cfg.tasks}
    for i in range(N // chunk):
        non_seq, seq, _ = O.synthetic_batch(ocfg, chunk, seq_lens, seed=500 + i)
        non_seq, seq = bf16_round_inputs(non_seq, seq)
        o = O.model_forward(P, ocfg, non_seq, seq)
        with torch.no_grad():
            g = model(to_cuda(non_seq), to_cuda(seq))
        for t in cfg.tasks:
            po[t].append(o[t].flatten())
            pg[t].append(g[t].flatten().float().cpu())
    gen = torch.Generator().manual_seed(1)
    for t in cfg.tasks:
        a, b = torch.cat(po[t]), torch.cat(pg[t])
        logit = torch.logit(a.clamp(1e-6, 1 - 1e-6)) * 3.0          # sharpen: random-init logits are small
        y = (torch.rand(N, generator=gen) < torch.sigmoid(logit)).numpy()
        auc_o, auc_g = roc_auc_score(y, a.numpy()), roc_auc_score(y, b.numpy())
        print(f'C1 AUC[{t}] oracle {auc_o:.6f} kernels {auc_g:.6f} delta {abs(auc_o - auc_g):.2e}')
        assert abs(auc_o - auc_g) <= 1e-4


def test_block_and_mha_accept_a_kv_cache_under_no_grad():
    """The reference's block-level ``kv_cache`` argument (OT/model.py:76,95-98,186): under ``torch.no_grad()`` the cached keys /
    values are placed in front; the result equals the same block on the concatenated sequence (queries = the new rows).  With
    gradients enabled it is refused (an inference feature)."""
    ocfg, cfg = make_configs(num_layers=1, num_ns_tokens=4, pyramid_enabled=False)
    P, model = _build(ocfg, cfg)
    blk = model.blocks[0]
    with torch.no_grad():     # every weight group = the shared one, so that a position's weights do not depend on the split point
        for w in (blk.attention.Wqkv, blk.ffn.W1, blk.ffn.b1, blk.ffn.W2, blk.ffn.b2):
            w.copy_(w[0:1].expand_as(w).clone())
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 24, 256, generator=g).to(torch.bfloat16).cuda()
    with torch.no_grad():
        y_full, (k_full, v_full) = blk(x)
        _, (k_old, v_old) = blk(x[:, :16])
        y_new, (k_cat, v_cat) = blk(x[:, 16:], kv_cache=(k_old, v_old))
        assert k_cat.shape == k_full.shape and rel_l2(k_cat, k_full) < 1e-2 and rel_l2(v_cat, v_full) < 1e-2
        assert rel_l2(y_new, y_full[:, 16:]) < 1e-2
        xn = x.clone()
        o_new, _ = blk.attention(xn[:, 16:], kv_cache=(k_old, v_old))
        assert o_new.shape == (4, 8, 256) and torch.isfinite(o_new.float()).all()
    with pytest.raises(RuntimeError, match='no_grad'):
        blk(x[:, 16:].requires_grad_(True), kv_cache=(k_old, v_old))
    with pytest.raises(RuntimeError, match='no_grad'):
        blk.attention(x[:, 16:].requires_grad_(True), kv_cache=(k_old, v_old))


def test_eval_forward_saves_no_activations():
    """an evaluation forward (``torch.no_grad()``) must not run in save mode - no GELU pre-activation output, no ctx."""
    from recommend_b200 import ops
    ocfg, cfg = make_configs(num_layers=1, num_ns_tokens=4, pyramid_enabled=False)
    P, model = _build(ocfg, cfg)
    non_seq, seq, _ = O.synthetic_batch(ocfg, 8, (20, 10, 5))
    seen = []
    real_gemm, real_ffn = ops.mixed_gemm, ops.ffn_fused

    def spy_gemm(*a, **kw):
        seen.append(kw.get('out2') is not None)
        return real_gemm(*a, **kw)

    def spy_ffn(*a, **kw):
        seen.append(kw.get('pre') is not None)
        return real_ffn(*a, **kw)
    from recommend_b200 import engine
    engine.ops.mixed_gemm, engine.ops.ffn_fused = spy_gemm, spy_ffn
    try:
        with torch.no_grad():
            model(to_cuda(non_seq), to_cuda(seq))
        assert seen and not any(seen)
        seen.clear()
        model(to_cuda(non_seq), to_cuda(seq), training=True)
        assert any(seen)
    finally:
        engine.ops.mixed_gemm, engine.ops.ffn_fused = real_gemm, real_ffn


def test_dropout_forward_backward_consistency():
    """training=True with the reference's dropout (OT/config.py:50, OT/model.py:184,193,198).  The mask is a counter-based
    hash, so parity with the oracle's torch RNG can only be statistical; what is checked exactly is that the forward
    epilogue mask and the backward mask kernel are the same function: extract the mask by pushing ones through the
    backward kernel, then compare one block's forward/backward with a torch autograd evaluation that uses that mask."""
    from recommend_b200 import ops, engine
    torch.manual_seed(0)
    rows, d = 700, 256
    ones = torch.ones(rows, d, dtype=torch.bfloat16, device='cuda')
    rate, seed = 0.1, 12345
    m = ops.dropout_mask(ones, seed, rate).float()
    keep = (m > 0).float()
    assert torch.allclose(m[m > 0], torch.full_like(m[m > 0], 1 / (1 - rate)), rtol=1e-2)       # inverted dropout scale
    assert abs(keep.mean().item() - (1 - rate)) < 5e-3                                          # drop fraction ~ rate
    assert abs(keep[:, ::2].mean().item() - keep[:, 1::2].mean().item()) < 1e-2                 # both hash lanes behave
    assert not torch.equal(keep, (ops.dropout_mask(ones, seed + 1, rate) > 0).float())          # seed matters
    # GEMM epilogue applies the same mask: out = res + mask * (A W^T)
    A = (torch.randn(rows, 64, device='cuda')).to(torch.bfloat16)
    W = (torch.randn(1, d, 64, device='cuda') * 0.2).to(torch.bfloat16)
    res = torch.randn(rows, d, device='cuda').to(torch.bfloat16)
    out = torch.empty(rows, d, dtype=torch.bfloat16, device='cuda')
    ops.mixed_gemm(A, W, [(0, 1, rows, 0, 0)], out, flags=4, res=res, dropout=(seed, rate))
    ref = res.float() + m * (A.float() @ W[0].float().t())
    assert ((out.float() - ref).abs() / (1 + ref.abs())).max().item() < 2e-2

    # whole model: training=True runs, is deterministic under torch.manual_seed, differs from eval, grads are finite
    ocfg, cfg = make_configs(num_layers=2, num_ns_tokens=8, schedule='linear_to_ns')
    cfg.dropout_rate = 0.1
    P, model = _build(ocfg, cfg)
    non_seq, seq, labels = O.synthetic_batch(ocfg, 64, (30, 20, 10))
    ns, sq = to_cuda(non_seq), to_cuda(seq)
    torch.manual_seed(7)
    a = model(ns, sq, training=True, return_logits=True)
    torch.manual_seed(7)
    b = model(ns, sq, training=True, return_logits=True)
    e = model(ns, sq, training=False, return_logits=True)
    assert all(torch.equal(a[t], b[t]) for t in cfg.tasks)
    assert not torch.equal(a['ctr'], e['ctr'])
    assert rel_l2(a['ctr'], e['ctr']) < 0.5     # dropout noise, same function underneath
    model.zero_grad(set_to_none=True)
    torch.manual_seed(7)
    out2 = model(ns, sq, training=True, return_logits=True)
    (out2['ctr'].sum() + out2['cvr'].sum()).backward()
    g = R.export_reference_style_params(model, grads=True)
    assert all(torch.isfinite(v).all() for v in g.values() if v is not None)
    assert g['blocks.0.ffn.W1'].abs().sum() > 0


def test_block_dropout_backward_matches_autograd_with_extracted_mask():
    """One OneTransBlock with dropout: forward and input gradient against torch autograd using the masks extracted from the
    backward mask kernel (the oracle block with identical weights, dropout applied explicitly)."""
    from recommend_b200 import ops
    ocfg, cfg = make_configs(num_layers=1, num_ns_tokens=4, pyramid_enabled=False)
    cfg.dropout_rate = 0.1
    P, model = _build(ocfg, cfg)
    blk = model.blocks[0]
    B, L, d = 16, 20, 256
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, L, d, generator=g).to(torch.bfloat16)
    xg = x.cuda().requires_grad_(True)
    torch.manual_seed(11)
    y, _ = blk(xg, training=True)
    dy = torch.randn(B, L, d, generator=g).to(torch.bfloat16)
    (y.float() * dy.cuda().float()).sum().backward()
    # reproduce the two seeds the block drew
    torch.manual_seed(11)
    s = torch.randint(0, 2 ** 31 - 1, (2,))
    ones = torch.ones(L * B, d, dtype=torch.bfloat16, device='cuda')
    m_att = ops.dropout_mask(ones, int(s[0]), 0.1).float().cpu().view(L, B, d).transpose(0, 1)
    m_ffn = ops.dropout_mask(ones, int(s[1]), 0.1).float().cpu().view(L, B, d).transpose(0, 1)
    xo = x.float().requires_grad_(True)
    b0 = 'blocks.0.'
    xn = O.rmsnorm(xo, P[b0 + 'norm1.scale'])
    z = xo + m_att * O.mixed_mha(P, b0 + 'attention.', ocfg, xn, L)
    zn = O.rmsnorm(z, P[b0 + 'norm2.scale'])
    yo = z + m_ffn * O.mixed_ffn(P, b0 + 'ffn.', ocfg, zn, 0, L)
    (yo * dy.float()).sum().backward()
    assert rel_l2(y.detach().float().cpu(), yo.detach()) < 1e-2
    assert rel_l2(xg.grad.float().cpu(), xo.grad) < 3e-2
