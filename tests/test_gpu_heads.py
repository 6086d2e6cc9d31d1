"""GPU parity of the fp32 head kernels (ot_heads_fwd / ot_heads_bwd: output RMSNorm, Dense(d/2, gelu), Dense(1, sigmoid), Keras BCE
- OT/model.py:322-330, 384-391; OT/train.py:84-87) against the same arithmetic in fp64 torch and against the oracle's BCE."""
import pytest
import torch

from oracle import onetrans_oracle as O
from recommend_b200 import ops

pytestmark = pytest.mark.gpu


def _ref(x, gain, eps, heads, labels):
    x = x.double().requires_grad_(True)
    gain = gain.double().requires_grad_(True)
    hs = [[w.double().requires_grad_(True) for w in h] for h in heads]
    xn = x * torch.rsqrt(x.square().mean(-1, keepdim=True) + eps) * gain
    logits, probs = [], []
    for k0, b0, k1, b1 in hs:
        h = torch.nn.functional.gelu(xn @ k0 + b0)          # exact erf form
        lg = (h @ k1 + b1).squeeze(-1)
        logits.append(lg); probs.append(torch.sigmoid(lg))
    logits, probs = torch.stack(logits), torch.stack(probs)
    loss = O.bce_loss({str(t): probs[t].unsqueeze(1) for t in range(len(hs))},
                      {str(t): labels[t].double().unsqueeze(1) for t in range(len(hs))}, [str(t) for t in range(len(hs))])
    return x, gain, hs, logits, probs, loss


@pytest.mark.parametrize('B,d,T', [(37, 256, 2), (2048, 256, 2), (50, 384, 1), (16, 128, 3)])
def test_heads_forward_backward_and_bce(B, d, T):
    torch.manual_seed(B + d)
    Hd = d // 2
    x = torch.randn(B, d, device='cuda') * 1.5
    gain = 1.0 + 0.1 * torch.randn(d, device='cuda')
    heads = [(torch.randn(d, Hd, device='cuda') * 0.08, torch.randn(Hd, device='cuda') * 0.1,
              torch.randn(Hd, 1, device='cuda') * 0.2, torch.randn(1, device='cuda') * 0.1) for _ in range(T)]
    labels = (torch.rand(T, B, device='cuda') < 0.5).float()
    probs, logits, loss, saved = ops.heads_fwd(x, gain, 1e-6, heads, labels)
    xr, gr, hr, logits_r, probs_r, loss_r = _ref(x.cpu(), gain.cpu(), 1e-6, [[w.cpu() for w in h] for h in heads], labels.cpu())
    assert (logits.cpu().double() - logits_r).abs().max().item() < 2e-5 * max(1.0, logits_r.abs().max().item())
    assert (probs.cpu().double() - probs_r).abs().max().item() < 1e-5
    assert abs(float(loss) - float(loss_r)) < 2e-5 * max(1.0, abs(float(loss_r)))
    # backward of the fused loss: dlogit = g_bce (upstream gradient 1)
    loss_r.backward()
    grads = [tuple(torch.zeros_like(w) for w in h) for h in heads]
    dgain = torch.zeros_like(gain)
    dx = ops.heads_bwd(x, gain, 1e-6, heads, saved, saved[3].clone(), grads, dgain)
    torch.cuda.synchronize()

    def close(a, b, what):
        a, b = a.cpu().double().reshape(b.shape), b
        assert (a - b).abs().max().item() <= 2e-4 * max(1e-6, b.abs().max().item()), what
    close(dx, xr.grad, 'dx')
    close(dgain, gr.grad, 'dgain')
    for t in range(T):
        for got, want, nm in zip(grads[t], hr[t], ('dW0', 'db0', 'dW1', 'db1')):
            close(got, want.grad, (nm, t))
    # no labels: no loss, same probabilities
    probs2, _, loss2, _ = ops.heads_fwd(x, gain, 1e-6, heads, None)
    assert loss2 is None and torch.equal(probs2, probs)


def test_heads_reject_unsupported_shapes():
    from recommend_b200._lib import OneTransLibraryError
    x = torch.randn(4, 96, device='cuda')
    heads = [(torch.randn(96, 48, device='cuda'), torch.zeros(48, device='cuda'), torch.randn(48, 1, device='cuda'), torch.zeros(1, device='cuda'))]
    with pytest.raises(OneTransLibraryError):          # 256 / 48 = 5 thread groups do not divide 16 samples
        ops.heads_fwd(x, torch.ones(96, device='cuda'), 1e-6, heads, None)
