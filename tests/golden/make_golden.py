"""Generates tests/golden/oracle_golden.json from the oracle (run: python -m tests.golden.make_golden).

The reference has no golden vectors and cannot be imported here, so these vectors only
pin the oracle against regressions and give the GPU tests a fixed target that does not depend on the
oracle code being unchanged."""
import json
import os

import torch

from oracle import onetrans_oracle as O


def cases():
    return {
        'smoke_main': dict(hidden_dim=256, num_layers=2, ffn_dim=512, num_ns_tokens=4, B=2, seq_lens=(10, 5, 7), seed=0),
        'tiny_tail': dict(hidden_dim=256, num_layers=3, ffn_dim=256, num_ns_tokens=4, B=3, seq_lens=(12, 9, 7), seed=1),
        'tiny_head_literal': dict(hidden_dim=256, num_layers=2, ffn_dim=256, num_ns_tokens=4, B=3, seq_lens=(12, 9, 7), seed=2,
                                  alignment='head_literal'),
    }


def run_case(spec):
    cfg = O.OracleConfig(hidden_dim=spec['hidden_dim'], num_layers=spec['num_layers'], num_heads=4, ffn_dim=spec['ffn_dim'],
                         num_ns_tokens=spec['num_ns_tokens'], ns_param_alignment=spec.get('alignment', 'tail'))
    P = O.init_params(cfg, seed=spec['seed'])
    O.randomize_small_params(P, seed=spec['seed'] + 1)
    non_seq, seq, labels = O.synthetic_batch(cfg, spec['B'], spec['seq_lens'], seed=1234 + spec['seed'])
    seq = {k: v.to(torch.bfloat16).float() for k, v in seq.items()}
    logits = O.model_forward(P, cfg, non_seq, seq, return_logits=True)
    loss, grads, _ = O.loss_and_grads(P, cfg, non_seq, seq, labels)
    out = {f'logits.{t}': v.flatten().tolist() for t, v in logits.items()}
    out['loss'] = [float(loss)]
    out['grad_norm.Wo0'] = [float(grads['blocks.0.attention.Wo'].norm())]
    out['grad_norm.W1_0'] = [float(grads['blocks.0.ffn.W1'].norm())]
    out['grad_norm.sep'] = [float(grads['tokenizer.sep_embedding'].norm())]
    return out


if __name__ == '__main__':
    res = {name: run_case(spec) for name, spec in cases().items()}
    path = os.path.join(os.path.dirname(__file__), 'oracle_golden.json')
    json.dump(res, open(path, 'w'), indent=1)
    print('wrote', path)
