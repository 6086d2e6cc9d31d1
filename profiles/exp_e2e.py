"""Where does the end-to-end arm lose time against the device-resident arm?  Four loops over the same C2 step:
device inputs / no sync, device inputs / loss read each step, host inputs copied in-line, host inputs prefetched."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import recommend_b200 as R
from recommend_b200.train import FlatGradBuffer, DevicePrefetcher, train_loop, train_step
from recommend_b200.data import create_sample_batch

B, K = 2048, 8
dev = torch.device('cuda', 0)
cfg = R.get_model_config('small'); cfg.num_ns_tokens = 32; cfg.pyramid_schedule = 'linear_to_ns'; cfg.dropout_rate = 0.1
torch.manual_seed(0)
model = R.OneTransModel(cfg).to(dev)
grads = FlatGradBuffer(model.parameters())
ns, sq, lb = create_sample_batch(cfg, B, (170, 170, 170))
h = ({k: v.pin_memory() for k, v in ns.items()}, {k: v.to(torch.bfloat16).pin_memory() for k, v in sq.items()}, {k: v.pin_memory() for k, v in lb.items()})
d = tuple({k: v.to(dev) for k, v in x.items()} for x in h)

def timed(fn):
    fn(2); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); fn(K); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, (time.perf_counter() - t0) * 1e3 / K

def dev_nosync(n):
    for _ in range(n): train_step(model, grads, *d)
def dev_sync(n):
    for _ in range(n): float(train_step(model, grads, *d))
def host_inline(n):
    for _ in range(n):
        x = tuple({k: v.to(dev, non_blocking=True) for k, v in t.items()} for t in h)
        float(train_step(model, grads, *x))
def host_prefetch(n):
    for x in DevicePrefetcher((h for _ in range(n)), dev):
        float(train_step(model, grads, *x))
def loop_api(n):
    train_loop(model, grads, (h for _ in range(n)), 1, None, dev)
def copy_only(n):
    for _ in range(n):
        x = tuple({k: v.to(dev, non_blocking=True) for k, v in t.items()} for t in h)
for name, fn in [('device inputs, no sync', dev_nosync), ('device inputs, loss read', dev_sync), ('host inputs in-line', host_inline),
                 ('host inputs prefetched', host_prefetch), ('train_loop (prefetch + lagged loss read)', loop_api), ('copies only', copy_only)]:
    ev, wall = timed(fn)
    print(f'{name:42s} {ev:7.2f} ms/step (events)  {wall:7.2f} ms/step (host clock)')
