"""GPU parity of the fused clip + RMSprop update (``ot_clip_rmsprop_step``) against the oracle's restatement of
OT/train.py:131-138.  fp32 arithmetic on both sides; tolerance 2e-6 relative (rsqrt / summation-order differences)."""
import pytest
import torch

from oracle import onetrans_oracle as O

pytestmark = pytest.mark.gpu


def _make(shapes, seed, scale):
    g = torch.Generator().manual_seed(seed)
    params = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]
    grads = [(torch.randn(*s, generator=g) * sc) for s, sc in zip(shapes, scale)]
    return params, grads


@pytest.mark.parametrize('momentum,clip', [(0.0, 0.0), (0.99999, 90.0), (0.5, 1.0)])
def test_clip_rmsprop_matches_oracle(momentum, clip):
    from recommend_b200.train import FlatGradBuffer, ClipRMSprop
    # ragged sizes: scalars, non-multiples of 4, exactly one chunk, > one chunk, a big matrix
    shapes = [(1,), (3,), (1024,), (1025,), (7, 33), (256, 1024), (5, 256, 768), (2,)]
    scale = [1.0, 100.0, 0.01, 5.0, 30.0, 0.5, 0.2, 1e-4]
    params, grads = _make(shapes, 7, scale)
    buf = FlatGradBuffer(params)
    opt = ClipRMSprop(buf, lr=0.005, rho=0.9, momentum=momentum, eps=1e-7, clip_norm=clip)
    ref_w = {i: p.detach().cpu().double() for i, p in enumerate(params)}
    ref_state = {}
    v0 = [p._version for p in params]
    for step in range(3):
        step_grads = {i: g.double() * (1.0 + 0.5 * step) for i, g in enumerate(grads)}
        for i, p in enumerate(params):
            p.grad.copy_(step_grads[i].float())
        opt.step()
        O.clip_rmsprop_update(ref_w, {i: g.float().double() for i, g in step_grads.items()}, ref_state, lr=0.005, rho=0.9,
                              momentum=momentum, eps=1e-7, clip_norm=clip)
        torch.cuda.synchronize()
        for i, p in enumerate(params):
            got, want = p.detach().cpu().double(), ref_w[i]
            err = (got - want).abs().max().item() / max(want.abs().max().item(), 1e-6)
            assert err < 2e-6, (step, i, shapes[i], err)
        if clip > 0:
            want_n = torch.tensor([step_grads[i].float().double().norm().item() for i in range(len(params))], dtype=torch.float64)
            assert torch.allclose(opt.grad_norms().cpu().double(), want_n, rtol=1e-5)
    assert all(p._version > v for p, v in zip(params, v0))     # compute copies will refresh
    # alignment padding between tensors is never written
    flat = buf.flat
    mask = torch.ones_like(flat, dtype=torch.bool)
    for p, o in zip(buf.params, buf.offsets):
        mask[o:o + p.numel()] = False
    assert float(opt.rms[mask].abs().sum()) == 0.0


def test_zero_grad_and_grad_scale():
    from recommend_b200.train import FlatGradBuffer, ClipRMSprop
    params, grads = _make([(130,), (64, 64)], 3, [1.0, 1.0])
    buf = FlatGradBuffer(params)
    opt = ClipRMSprop(buf, lr=0.01, momentum=0.0, clip_norm=0.5)
    w0 = [p.detach().clone() for p in params]
    for p, g in zip(params, grads):
        p.grad.copy_(g * 4.0)
    opt.step(grad_scale=0.25, zero_grad=True)
    torch.cuda.synchronize()
    assert float(buf.flat.abs().sum()) == 0.0
    for p, g, w in zip(params, grads, w0):
        gc = O.clip_by_norm(g.double(), 0.5)
        want, _, _ = O.rmsprop_step(w.cpu().double(), gc, torch.zeros_like(gc), None, 0.01, 0.9, 0.0, 1e-7)
        assert (p.detach().cpu().double() - want).abs().max().item() < 2e-6


def test_train_step_with_optimizer_reduces_loss():
    """End to end: a few clip+RMSprop steps on one fixed batch lower the BCE (OT/train.py:116-138)."""
    import recommend_b200 as R
    from recommend_b200.train import FlatGradBuffer, ClipRMSprop, train_step
    cfg = R.get_model_config('small')
    cfg.num_layers, cfg.num_ns_tokens, cfg.dropout_rate = 2, 4, 0.0
    torch.manual_seed(0)
    model = R.OneTransModel(cfg).cuda()
    ocfg = O.small_config(num_ns_tokens=4)
    non_seq, seq, labels = O.synthetic_batch(ocfg, 64, (20, 20, 20))
    non_seq = {k: v.cuda() for k, v in non_seq.items()}
    seq = {k: v.cuda() for k, v in seq.items()}
    labels = {k: v.cuda() for k, v in labels.items()}
    buf = FlatGradBuffer(model.parameters())
    opt = ClipRMSprop(buf, lr=1e-3, rho=0.9, momentum=0.0, clip_norm=cfg.gradient_clip_norm)
    losses = [float(train_step(model, buf, non_seq, seq, labels, optimizer=opt)) for _ in range(8)]
    assert losses[-1] < losses[0], losses


@pytest.mark.parametrize('d', [128, 64])
def test_clip_granularity_is_the_keras_variable(d):
    """OT/train.py:135 clips every entry of ``model.trainable_variables``: each (weight group, q | k | v) Dense kernel and each
    per-group FFN kernel / bias on its own, not the packed ``[G, ...]`` tensors as wholes.  d = 64 makes the
    q | k | v column blocks narrower than a warp's 128 elements (the kernel's per-lane path)."""
    from recommend_b200.train import FlatGradBuffer, ClipRMSprop
    G, Fd = 3, 2 * d
    g = torch.Generator().manual_seed(11)
    mk = lambda *s: torch.nn.Parameter(torch.randn(*s, generator=g).cuda())
    Wqkv, W1, b1, plain = mk(G, d, 3 * d), mk(G, d, Fd), mk(G, Fd), mk(d, d)
    Wqkv._ot_clip_layout = (d * 3 * d, 3 * d, d)
    W1._ot_clip_layout = (d * Fd,) * 3
    b1._ot_clip_layout = (Fd,) * 3
    params = [Wqkv, W1, b1, plain]
    buf = FlatGradBuffer(params)
    clip = 2.0
    opt = ClipRMSprop(buf, lr=0.01, momentum=0.0, clip_norm=clip)
    assert opt.n_slots == 3 * G + G + G + 1
    w0 = [p.detach().cpu().double() for p in params]
    # gradients whose per-variable norms straddle the clip: some variables are clipped, some are not
    grads = [torch.randn(p.shape, generator=g, dtype=torch.float64) * sc for p, sc in zip(params, (0.02, 0.005, 0.3, 0.01))]
    grads[0][1, :, d:2 * d] *= 20.0            # only the k kernel of group 1 is far above the clip
    grads[1][2] *= 30.0
    for p, gr in zip(params, grads):
        p.grad.copy_(gr.float())
    opt.step()
    torch.cuda.synchronize()

    def clipped(i, gr):
        gr = gr.float().double()
        if i == 0:
            out = gr.clone()
            for gi in range(G):
                for part in range(3):
                    sl = (gi, slice(None), slice(part * d, (part + 1) * d))
                    out[sl] = O.clip_by_norm(gr[sl], clip)
            return out
        if i in (1, 2):
            return torch.stack([O.clip_by_norm(gr[gi], clip) for gi in range(G)])
        return O.clip_by_norm(gr, clip)

    n_clipped = 0
    for i, (p, gr, w) in enumerate(zip(params, grads, w0)):
        gc = clipped(i, gr)
        n_clipped += int(not torch.equal(gc, gr.float().double()))
        want, _, _ = O.rmsprop_step(w, gc, torch.zeros_like(gc), None, 0.01, 0.9, 0.0, 1e-7)
        err = (p.detach().cpu().double() - want).abs().max().item() / want.abs().max().item()
        assert err < 2e-6, (i, err)
    assert n_clipped >= 2
    # the packed-tensor granularity of round 1 gives a different update: the test can tell the two apart
    whole = O.clip_by_norm(grads[0].float().double(), clip)
    assert not torch.allclose(whole, clipped(0, grads[0]), rtol=1e-3)
    norms = opt.grad_norms().cpu().double()
    assert torch.allclose(norms[1 * 3 + 1], grads[0][1, :, d:2 * d].float().double().norm(), rtol=1e-5)
