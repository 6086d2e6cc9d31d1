"""CPU experiment (no GPU): which bf16 rounding sites of the CUDA pipeline cost how much logits error?

Emulates the data path of recommend_b200/engine.py in fp32 on the CPU with a switchable bf16 rounding at every place the
kernels store an activation, separately for sequence (S) rows and non-sequence (NS) rows, and reports the
logits relative L2 error against the un-rounded oracle (oracle/onetrans_oracle.py).  Round-2 VERDICT item 1(b).

    python profiles/exp_rounding_ablation.py [--layers 6] [--B 32]
"""
import argparse, math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import onetrans_oracle as O

bf = lambda t: t.to(torch.bfloat16).to(torch.float32)


def rnd(t, ns0, on_s, on_ns):
    """round rows [:, :ns0] if on_s, rows [:, ns0:] if on_ns (t: [B, n, *], NS tokens are the last rows)"""
    if not on_s and not on_ns:
        return t
    out = t.clone()
    if on_s and ns0 > 0:
        out[:, :ns0] = bf(t[:, :ns0])
    if on_ns:
        out[:, ns0:] = bf(t[:, ns0:])
    return out


def forward(P, cfg, non_seq, seq, sites):
    """sites: set of strings '<site>_<s|ns>' with site in resid, xn, qkv, p, o, zn, h"""
    S = lambda name: (f'{name}_s' in sites, f'{name}_ns' in sites)
    x = O.tokenizer_forward(P, cfg, non_seq, seq)
    L0 = x.shape[1]
    L_ns = cfg.num_ns_tokens
    keep_lens = O.resolve_keep_lens(cfg, L0)
    H, d = cfg.num_heads, cfg.hidden_dim
    dh = d // H
    al = cfg.ns_param_alignment
    x = rnd(x, L0 - L_ns, *S('resid'))
    for l in range(cfg.num_layers):
        b = f'blocks.{l}.'
        cur, keep = x.shape[1], keep_lens[l]
        ns0 = cur - min(L_ns, cur)
        xn = rnd(O.rmsnorm(x, P[b + 'norm1.scale']), ns0, *S('xn'))
        pa = b + 'attention.'
        k = rnd(O._mixed_linear(xn, P[pa + 'Wk'], None, 0, cur, L_ns, al, False), ns0, *S('qkv'))
        v = rnd(O._mixed_linear(xn, P[pa + 'Wv'], None, 0, cur, L_ns, al, False), ns0, *S('qkv'))
        tns0 = max(ns0 - (cur - keep), 0)     # first NS row inside the tail
        q = rnd(O._mixed_linear(xn[:, cur - keep:], P[pa + 'Wq'], None, cur - keep, cur, L_ns, al, False), tns0, *S('qkv'))
        B = x.shape[0]
        q4, k4, v4 = q.reshape(B, keep, H, dh), k.reshape(B, cur, H, dh), v.reshape(B, cur, H, dh)
        sc = torch.einsum('bqhd,bkhd->bhqk', q4, k4) / math.sqrt(float(dh))
        qi = torch.arange(keep).unsqueeze(1) + (cur - keep)
        ki = torch.arange(cur).unsqueeze(0)
        sc = torch.where(ki <= qi, sc, torch.full_like(sc, -1e9))
        m = sc.max(dim=-1, keepdim=True).values
        p = torch.exp(sc - m)
        l_sum = p.sum(dim=-1, keepdim=True)
        ps, pn = S('p')
        if ps or pn:   # kernel: unnormalised P rounded to bf16 before the PV product, row sum from the fp32 values
            pr = p.permute(0, 2, 1, 3)        # [B, q, H, k]
            pr = rnd(pr, tns0, ps, pn).permute(0, 2, 1, 3)
        else:
            pr = p
        o = (torch.einsum('bhqk,bkhd->bqhd', pr, v4) / l_sum.permute(0, 2, 1, 3)).reshape(B, keep, d)
        o = rnd(o, tns0, *S('o'))
        z = x[:, cur - keep:] + o @ P[pa + 'Wo']
        z = rnd(z, tns0, *S('resid'))
        zn = rnd(O.rmsnorm(z, P[b + 'norm2.scale']), tns0, *S('zn'))
        pf = b + 'ffn.'
        h = O.gelu_erf(O._mixed_linear(zn, P[pf + 'W1'], P[pf + 'b1'], cur - keep, cur, L_ns, al, False))
        h = rnd(h, tns0, *S('h'))
        y = z + O._mixed_linear(h, P[pf + 'W2'], P[pf + 'b2'], cur - keep, cur, L_ns, al, False)
        x = rnd(y, tns0, *S('resid'))
    xo = O.rmsnorm(x, P['output_norm.scale'])
    lg = O.heads_forward(P, cfg, xo[:, -1, :])
    return torch.cat([lg[t].flatten() for t in cfg.tasks])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--layers', type=int, default=6)
    ap.add_argument('--B', type=int, default=32)
    ap.add_argument('--ns', type=int, default=16)
    ap.add_argument('--seed', type=int, default=0)
    a = ap.parse_args()
    cfg = O.OracleConfig(hidden_dim=256, num_layers=a.layers, num_heads=4, ffn_dim=1024, num_ns_tokens=a.ns, dropout_rate=0.0)
    P = O.init_params(cfg, seed=a.seed)
    O.randomize_small_params(P, seed=a.seed + 1)
    for k in P:
        if P[k].dim() >= 2 and 'ns_tokenizer' not in k and 'task_heads' not in k and 'sep_embedding' not in k:
            P[k] = bf(P[k])
    non_seq, seq, _ = O.synthetic_batch(cfg, a.B, (86, 84, 84), seed=1234 + a.seed)
    seq = {k: bf(v) for k, v in seq.items()}
    L0 = 256 + a.ns
    cfg.pyramid_keep_lens = O.keep_lens_linear_to_ns(L0, a.layers, a.ns)
    ref = forward(P, cfg, non_seq, seq, set())
    chk = torch.cat([v.flatten() for v in O.model_forward(P, cfg, non_seq, seq, return_logits=True).values()])
    print('emulation == oracle without rounding:', float((ref - chk).norm() / chk.norm()))
    e = lambda s: float((forward(P, cfg, non_seq, seq, set(s)) - ref).norm() / ref.norm())
    names = ['resid', 'xn', 'qkv', 'p', 'o', 'zn', 'h']
    current = [f'{n}_s' for n in names] + [f'{n}_ns' for n in names if n != 'resid']
    print(f'current pipeline (all sites, NS residual fp32): {e(current):.3e}')
    print(f'everything bf16 (no fp32 NS stream):            {e(current + ["resid_ns"]):.3e}')
    for n in names:
        for side in ('s', 'ns'):
            print(f'  only {n}_{side:2s}: {e([f"{n}_{side}"]):.3e}     current minus it: {e([c for c in current if c != f"{n}_{side}"]):.3e}')
    print(f'no S-side rounding at all : {e([c for c in current if c.endswith("_ns")]):.3e}')
    print(f'no NS-side rounding at all: {e([c for c in current if c.endswith("_s")]):.3e}')
    print(f'NS side without xn, zn    : {e([c for c in current if c not in ("xn_ns", "zn_ns")]):.3e}')
    print(f'NS side without xn, zn, h : {e([c for c in current if c not in ("xn_ns", "zn_ns", "h_ns")]):.3e}')
    print(f'NS side without xn, zn, h, o : {e([c for c in current if c not in ("xn_ns", "zn_ns", "h_ns", "o_ns")]):.3e}')


if __name__ == '__main__':
    main()
