"""Host-side cost of one train step: run C2 with a tiny batch so that the GPU is never the bottleneck, time the
step on the host clock and print the cProfile top of the Python/ctypes launch path."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import recommend_b200 as R
from recommend_b200.train import FlatGradBuffer, train_step
from recommend_b200.data import create_sample_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = R.get_model_config('small'); cfg.num_ns_tokens = 32; cfg.pyramid_schedule = 'linear_to_ns'; cfg.dropout_rate = 0.1
torch.manual_seed(0)
model = R.OneTransModel(cfg).cuda()
grads = FlatGradBuffer(model.parameters())
ns, sq, lb = create_sample_batch(cfg, B, (170, 170, 170))
ns = {k: v.cuda() for k, v in ns.items()}; sq = {k: v.cuda().bfloat16() for k, v in sq.items()}; lb = {k: v.cuda() for k, v in lb.items()}
for _ in range(3):
    train_step(model, grads, ns, sq, lb)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    train_step(model, grads, ns, sq, lb)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'B={B}: host {1e3 * (t1 - t0) / 10:.2f} ms/step issue time, {1e3 * (t2 - t0) / 10:.2f} ms/step incl. drain')
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    train_step(model, grads, ns, sq, lb)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
