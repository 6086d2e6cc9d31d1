"""ID front end at BASELINE config 2 size: 2048 samples x 3 sequences x 170 events, 4 fields x 16-d tables, 1 M rows per
table, Zipf(1.05) ids.  Prints algorithmic GB/s of the gather, the gradient scatter and the sparse
Adagrad step against the measured HBM copy peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import recommend_b200 as R

B, L, NSEQ, vocab, ef = 2048, 170, 3, [1_000_000] * 4, 16
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else 6650.0
torch.manual_seed(0)
emb = R.EventEmbedding(vocab, ef).cuda()
opt = R.SparseAdagrad(emb)
g = torch.Generator(device='cuda').manual_seed(1)
r = torch.arange(1, vocab[0] + 1, dtype=torch.float64, device='cuda').pow(-1.05)
ids = torch.multinomial(r / r.sum(), B * L * NSEQ * 4, replacement=True, generator=g).to(torch.int32).reshape(B, L * NSEQ, 4)
d = torch.randn(B, L * NSEQ, 4 * ef, device='cuda').to(torch.bfloat16)
n_ev = B * L * NSEQ


def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


ev = emb(ids)
t_g = timed(lambda: emb(ids))
def bwd():
    emb._touched.clear()
    emb(ids).backward(d)
t_gb = timed(bwd)
def step():
    emb._touched[:] = [ids.contiguous()]
    opt.step()
t_s = timed(step)
uniq = int(torch.unique(ids.reshape(-1, 4)[:, 0]).numel())
by_g = n_ev * (16 + 4 * ef * 4 + 4 * ef * 2)              # ids + fp32 rows + bf16 event
by_b = n_ev * (16 + 4 * ef * 2 + 4 * ef * 4 * 2)          # ids + bf16 event gradient + fp32 row read-modify-write
print(json.dumps({'events': n_ev, 'unique_rows_field0': uniq,
                  'gather_us': t_g, 'gather_gbs': by_g / t_g / 1e3, 'gather_frac_of_hbm': by_g / t_g / 1e3 / peak,
                  'scatter_us': t_gb - t_g, 'scatter_gbs': by_b / max(t_gb - t_g, 1e-3) / 1e3,
                  'adagrad_us': t_s, 'hbm_peak_gbs': peak}))
