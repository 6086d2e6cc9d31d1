"""GPU parity of the ID front end (ot_embed_gather_fwd / scatter_bwd / adagrad_step) against the oracle's restatement:
the lookup is a copy (bit-exact), the sparse gradient and the Adagrad update are fp32 (tolerance 1e-5 relative)."""
import pytest
import torch

from oracle import onetrans_oracle as O
import recommend_b200 as R

pytestmark = pytest.mark.gpu


def _zipf_ids(B, L, vocab_sizes, seed, a=1.05):
    g = torch.Generator().manual_seed(seed)
    cols = []
    for v in vocab_sizes:
        r = torch.arange(1, v + 1, dtype=torch.float64)
        p = r.pow(-a)
        cols.append(torch.multinomial(p / p.sum(), B * L, replacement=True, generator=g).reshape(B, L))
    return torch.stack(cols, dim=-1).to(torch.int32)


@pytest.mark.parametrize('B,L,vocab,ef', [(3, 5, [7, 11, 2, 30], 16), (64, 170, [1000, 5000, 50, 20000], 16), (2, 9, [13, 4], 8)])
def test_gather_is_a_bit_exact_copy_and_scatter_sums_duplicates(B, L, vocab, ef):
    torch.manual_seed(0)
    emb = R.EventEmbedding(vocab, ef).cuda()
    ids = _zipf_ids(B, L, vocab, 1)
    ev = emb(ids.cuda())
    want = O.embed_events(emb.table.cpu(), vocab, ids)
    assert ev.dtype == torch.bfloat16 and ev.shape == (B, L, len(vocab) * ef)
    assert torch.equal(ev.cpu(), want)                                   # bit-exact
    d = torch.randn(B, L, len(vocab) * ef).to(torch.bfloat16)
    ev.backward(d.cuda())
    g_want = O.embed_grad_table(emb.table.cpu(), vocab, ids, d.float())
    g_got = emb.grad_table.cpu().double()
    assert (g_got - g_want).abs().max().item() <= 1e-5 * max(1.0, g_want.abs().max().item())
    assert emb.out_of_vocabulary_count() == 0


def test_out_of_vocabulary_ids_give_zero_rows_and_are_counted():
    emb = R.EventEmbedding([4, 4], 16).cuda()
    ids = torch.tensor([[[1, 7], [-1, 2]]], dtype=torch.int32)
    ev = emb(ids.cuda()).float().cpu()
    assert ev[0, 0, 16:].abs().sum() == 0 and ev[0, 1, :16].abs().sum() == 0
    assert ev[0, 0, :16].abs().sum() > 0 and ev[0, 1, 16:].abs().sum() > 0
    assert emb.out_of_vocabulary_count() == 2


def test_sparse_adagrad_matches_dense_keras_rule():
    torch.manual_seed(1)
    vocab, ef = [50, 300, 9], 16
    emb = R.EventEmbedding(vocab, ef).cuda()
    opt = R.SparseAdagrad(emb, lr=0.1, initial_accumulator_value=0.1, eps=1e-7)
    w_ref, acc_ref = emb.table.cpu().double(), torch.full(emb.table.shape, 0.1, dtype=torch.float64)
    for step in range(3):
        batches = [_zipf_ids(8, 20, vocab, 10 * step + j) for j in range(2)]          # two sequences share the tables
        ds = [torch.randn(8, 20, len(vocab) * ef).to(torch.bfloat16) for _ in batches]
        g_ref = torch.zeros_like(w_ref)
        for ids, d in zip(batches, ds):
            emb(ids.cuda()).backward(d.cuda())
            g_ref += O.embed_grad_table(w_ref.float(), vocab, ids, d.float())
        opt.step()
        w_ref, acc_ref = O.adagrad_step(w_ref, g_ref, acc_ref, 0.1, 1e-7)
        torch.cuda.synchronize()
        assert (emb.table.cpu().double() - w_ref).abs().max().item() < 2e-6
        assert (opt.acc.cpu().double() - acc_ref).abs().max().item() < 1e-5 * acc_ref.abs().max().item()
        assert float(emb.grad_table.abs().sum()) == 0.0                           # every touched row was reset


def test_ids_to_loss_end_to_end():
    """ids -> EventEmbedding -> OneTransModel -> BCE -> backward -> dense RMSprop + sparse Adagrad: the loss goes down and
    only looked-up rows of the tables move."""
    from recommend_b200.train import FlatGradBuffer, ClipRMSprop, bce_loss
    cfg = R.get_model_config('small')
    cfg.num_layers, cfg.num_ns_tokens, cfg.dropout_rate = 2, 4, 0.0
    torch.manual_seed(0)
    model = R.OneTransModel(cfg).cuda()
    vocab = [100, 1000, 20, 500]
    emb = R.EventEmbedding(vocab, 16).cuda()
    ocfg = O.small_config(num_ns_tokens=4)
    non_seq, _, labels = O.synthetic_batch(ocfg, 32, (12, 12, 12))
    non_seq = {k: v.cuda() for k, v in non_seq.items()}
    labels = {k: v.cuda() for k, v in labels.items()}
    ids = {n: _zipf_ids(32, 12, vocab, 5 + i).cuda() for i, n in enumerate(cfg.feature_config['sequence_features'])}
    grads = FlatGradBuffer(model.parameters())
    dense, sparse = ClipRMSprop(grads, lr=1e-3, clip_norm=cfg.gradient_clip_norm), R.SparseAdagrad(emb, lr=0.05)
    t0 = emb.table.clone()
    losses = []
    for _ in range(6):
        grads.zero()
        seq = {n: emb(i) for n, i in ids.items()}
        loss = bce_loss(model(non_seq, seq, training=True), labels, cfg.tasks)
        loss.backward()
        dense.step(); sparse.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses
    moved = (emb.table - t0).abs().sum(dim=1) > 0
    touched = torch.zeros_like(moved)
    off = 0
    for f, v in enumerate(vocab):
        for i in ids.values():
            touched[off + i[..., f].long().flatten()] = True
        off += v
    assert torch.equal(moved & ~touched, torch.zeros_like(moved))       # untouched rows never move
    assert moved.any()
