import torch, sys
sys.path.insert(0, ".")
from recommend_b200 import ops
bf16=torch.bfloat16
import os
for (Lq,Lk) in [(458,544),(373,458),(288,373),(202,288),(117,202),(32,117)]:
    B,H,dh=2048,4,64; d=H*dh
    g=torch.Generator(device="cuda").manual_seed(0)
    q=torch.randn(Lq*B,d,generator=g,device="cuda").to(bf16); kv=torch.randn(Lk*B,2*d,generator=g,device="cuda").to(bf16)
    o=torch.empty(Lq*B,d,dtype=bf16,device="cuda"); lse=torch.empty(B*H*Lq,device="cuda")
    for _ in range(2): ops.attn_fwd(q,kv[:,:d],kv[:,d:],o,lse,B,H,Lq,Lk,dh)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.attn_fwd(q,kv[:,:d],kv[:,d:],o,lse,B,H,Lq,Lk,dh)
    e1.record(); torch.cuda.synchronize()
    print(Lq,Lk,"attn_fwd ms", round(e0.elapsed_time(e1)/5,4), flush=True)
