"""Golden vectors from the REFERENCE ITSELF: imports /root/reference/rank/scaling_up/oneTrans/practice/{config,model}.py unmodified,
with ``oracle/tf_shim.py`` standing in for the TensorFlow ops they call, runs the reference's own classes and writes
``tests/golden/reference_golden.npz`` (+ ``reference_golden.json`` for the integer / configuration facts).

    python -m tests.golden.make_reference_golden          # in the build container, where /root/reference exists

What runs as written (and is recorded):
  * ``OneTransConfig`` / ``OneTransSmallConfig`` / ``OneTransLargeConfig`` / ``get_model_config`` defaults (pure Python);
  * ``PyramidScheduler.get_layer_config`` for a sweep of lengths (pure Python, OT/model.py:287-302);
  * case A - ``OneTransModel.call`` with the pyramid off, 2 blocks: tokenizer -> blocks -> output norm -> heads (OT/model.py:335-393);
  * case B - ``OneTransModel.call`` with the pyramid ON and ONE block: the literal compute-every-query-then-gather path (:356-371);
  * case C - ``OneTransBlock`` alone on a random input, training=False, and its returned (k, v);
  * ``SequenceProcessor.process_sequence`` of OT/data_loader.py:68-101 (numpy: keep the newest events, pad in front) on an empty, a
    short and a long sequence;
  * the per-position weight rule as the code applies it (``_get_projection_weights``: position < num_ns_tokens -> dedicated), i.e. the
    literal 'head' alignment of SURVEY.md D4 - the oracle's ``ns_param_alignment='head_literal'`` mode is what these vectors pin.
What does NOT run as written (recorded as the exception the reference raises - SURVEY.md §A.3):
  * D2 - pyramid on with two blocks: layer 1 gathers indices of the ORIGINAL length from the already shortened sequence;
  * D9 - integer ids next to float features in ``tf.concat``;
  * D6 - the KV-cache branch: a second call with ``use_kv_cache=True`` puts the cached keys in front of the new ones while the mask
    keeps its built ``[L, L]`` shape;
  * D7 - the one-tuple call form of OT/train.py:118 / OT/evaluate.py:87 against ``call(non_seq_features, seq_features, ...)``;
  * D8 - ``config.gradient_clip`` (OT/train.py:134) does not exist, ``gradient_clip_norm`` does.
Every parameter of the reference model is exported under the oracle's names so that the oracle can be run on identical weights."""
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

from oracle import tf_shim

REF = '/root/reference/rank/scaling_up/oneTrans/practice'
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    tf_shim.install()
    pkg = types.ModuleType('ot_ref')
    pkg.__path__ = [REF]                      # a package without running the reference's __init__ (it imports pandas / matplotlib users)
    sys.modules['ot_ref'] = pkg
    return importlib.import_module('ot_ref.config'), importlib.import_module('ot_ref.model')


def export_params(model, cfg):
    """Reference model -> the oracle's parameter names (packed: index 0 shared, 1 + j dedicated j)."""
    P = {}
    tok = model.tokenizer
    if tok.ns_tokenizer.layers[0].kernel is not None:      # never built when no configured feature was present
        P['tokenizer.ns_tokenizer.kernel'] = tok.ns_tokenizer.layers[0].kernel
        P['tokenizer.ns_tokenizer.bias'] = tok.ns_tokenizer.layers[0].bias
    for i, dense in enumerate(tok.seq_projections):
        if dense.kernel is not None:          # a projection whose sequence never appeared is never built (lazy Keras build)
            P[f'tokenizer.seq_projections.{i}.kernel'], P[f'tokenizer.seq_projections.{i}.bias'] = dense.kernel, dense.bias
    if tok.sep_embedding.embeddings is not None:           # never built when no [SEP] was placed (only the last sequence present)
        P['tokenizer.sep_embedding'] = tok.sep_embedding.embeddings
    for l, blk in enumerate(model.blocks):
        P.update(export_block(blk, f'blocks.{l}.'))
    P['output_norm.scale'] = model.output_norm.scale
    for t, head in model.task_heads.items():
        for i, dense in enumerate(head.layers):
            P[f'task_heads.{t}.{i}.kernel'], P[f'task_heads.{t}.{i}.bias'] = dense.kernel, dense.bias
    return P


def export_block(blk, b):
    att, ffn = blk.attention, blk.ffn
    d = att.hidden_dim

    def stack(shared, dedicated, attr, shape):
        # positions the run never reached leave their Dense unbuilt: zeros there (the oracle never reads them either)
        return torch.stack([getattr(m, attr) if getattr(m, attr) is not None else torch.zeros(shape, dtype=tf_shim.FLOAT) for m in [shared] + list(dedicated)])
    P = {b + 'norm1.scale': blk.norm1.scale, b + 'norm2.scale': blk.norm2.scale, b + 'attention.Wo': att.Wo.kernel}
    P[b + 'attention.Wq'] = stack(att.Wq_shared, att.Wq_dedicated, 'kernel', (d, d))
    P[b + 'attention.Wk'] = stack(att.Wk_shared, att.Wk_dedicated, 'kernel', (d, d))
    P[b + 'attention.Wv'] = stack(att.Wv_shared, att.Wv_dedicated, 'kernel', (d, d))
    F = ffn.ffn_dim
    first = lambda seq: seq.layers[0]
    second = lambda seq: seq.layers[1]
    P[b + 'ffn.W1'] = stack(first(ffn.ffn_shared), [first(s) for s in ffn.ffn_dedicated], 'kernel', (d, F))
    P[b + 'ffn.b1'] = stack(first(ffn.ffn_shared), [first(s) for s in ffn.ffn_dedicated], 'bias', (F,))
    P[b + 'ffn.W2'] = stack(second(ffn.ffn_shared), [second(s) for s in ffn.ffn_dedicated], 'kernel', (F, d))
    P[b + 'ffn.b2'] = stack(second(ffn.ffn_shared), [second(s) for s in ffn.ffn_dedicated], 'bias', (d,))
    return P


def assign_params(model, P):
    """The inverse of ``export_params``: copy an oracle-named parameter dict into the LIVE tensors of a built reference model."""
    def put(dst, src):
        if dst is not None:
            dst.copy_(src.to(dst.dtype).reshape(dst.shape))
    tok = model.tokenizer
    put(tok.ns_tokenizer.layers[0].kernel, P['tokenizer.ns_tokenizer.kernel'])
    put(tok.ns_tokenizer.layers[0].bias, P['tokenizer.ns_tokenizer.bias'])
    for i, dense in enumerate(tok.seq_projections):
        put(dense.kernel, P[f'tokenizer.seq_projections.{i}.kernel'])
        put(dense.bias, P[f'tokenizer.seq_projections.{i}.bias'])
    put(tok.sep_embedding.embeddings, P['tokenizer.sep_embedding'])
    for l, blk in enumerate(model.blocks):
        b = f'blocks.{l}.'
        att, ffn = blk.attention, blk.ffn
        put(blk.norm1.scale, P[b + 'norm1.scale'])
        put(blk.norm2.scale, P[b + 'norm2.scale'])
        put(att.Wo.kernel, P[b + 'attention.Wo'])
        for name, shared, dedicated in (('Wq', att.Wq_shared, att.Wq_dedicated), ('Wk', att.Wk_shared, att.Wk_dedicated), ('Wv', att.Wv_shared, att.Wv_dedicated)):
            for g, dense in enumerate([shared] + list(dedicated)):
                put(dense.kernel, P[b + 'attention.' + name][g])
        for g, seq in enumerate([ffn.ffn_shared] + list(ffn.ffn_dedicated)):
            put(seq.layers[0].kernel, P[b + 'ffn.W1'][g])
            put(seq.layers[0].bias, P[b + 'ffn.b1'][g])
            put(seq.layers[1].kernel, P[b + 'ffn.W2'][g])
            put(seq.layers[1].bias, P[b + 'ffn.b2'][g])
    put(model.output_norm.scale, P['output_norm.scale'])
    for t, head in model.task_heads.items():
        for i, dense in enumerate(head.layers):
            put(dense.kernel, P[f'task_heads.{t}.{i}.kernel'])
            put(dense.bias, P[f'task_heads.{t}.{i}.bias'])


def run_kernel_shape_case(C, M, name, spec):
    """A case at a shape the sm_100a kernels support (d 256, head_dim 64), for the DIRECT product-vs-reference test on the GPU
    (tests/test_gpu_model.py::test_product_equals_the_reference_outputs).  Weights and inputs are not stored: both sides rebuild them
    with the recipe of tests/helpers.reference_case_inputs (oracle init_params seed + bf16 rounding, synthetic_batch seed); only the
    reference's outputs and a checksum of what went in are committed."""
    from tests.helpers import reference_case_inputs
    ocfg, P, non_seq, seq = reference_case_inputs(spec)
    cfg = C.OneTransConfig()
    cfg.hidden_dim, cfg.num_heads, cfg.ffn_dim, cfg.num_layers, cfg.num_ns_tokens = spec['hidden_dim'], spec['num_heads'], spec['ffn_dim'], spec['num_layers'], spec['num_ns_tokens']
    cfg.pyramid_enabled = spec['pyramid_enabled']
    tf_shim.set_seed(0)
    model = M.OneTransModel(cfg)
    f64 = lambda d: {k: v.to(tf_shim.FLOAT) for k, v in d.items()}
    model(f64(non_seq), f64(seq), training=False)            # builds the weights
    assign_params(model, P)
    model.reset_kv_cache()
    out = model(f64(non_seq), f64(seq), training=False)
    arrays = {f'{name}/out/prob/{t}': v for t, v in out.items()}
    checksum = float(sum(v.double().sum() for v in P.values()) + sum(v.double().sum() for v in non_seq.values()) + sum(v.double().sum() for v in seq.values()))
    return arrays, dict(spec, checksum=checksum, total_len=int(model.tokenizer(f64(non_seq), f64(seq)).shape[1]))


def perturb(root, seed):
    """Biases and norm gains of the LIVE reference layers away from their zero / one initial values (in place, before the export), so
    that the vectors exercise them.  Walks the layer tree the way it was built: attributes, lists, dicts."""
    g = torch.Generator().manual_seed(seed)
    seen = set()

    def visit(obj):
        if id(obj) in seen:
            return
        seen.add(id(obj))
        if isinstance(obj, tf_shim.Dense):
            if obj.bias is not None:
                obj.bias.copy_((torch.rand(obj.bias.shape, generator=g, dtype=obj.bias.dtype) * 2 - 1) * 0.1)
        elif isinstance(obj, tf_shim.Layer):
            if isinstance(getattr(obj, 'scale', None), torch.Tensor):           # RMSNorm (OT/model.py:14-17)
                obj.scale.copy_(1.0 + (torch.rand(obj.scale.shape, generator=g, dtype=obj.scale.dtype) * 2 - 1) * 0.1)
            for v in vars(obj).values():
                visit(v)
        elif isinstance(obj, (list, tuple)):
            for v in obj:
                visit(v)
        elif isinstance(obj, dict):
            for v in obj.values():
                visit(v)
    visit(root)


def small_config(C, num_layers, pyramid):
    cfg = C.OneTransConfig()
    cfg.hidden_dim, cfg.num_heads, cfg.ffn_dim, cfg.num_layers, cfg.num_ns_tokens = 16, 4, 24, num_layers, 3
    cfg.pyramid_enabled, cfg.dropout_rate = pyramid, 0.1
    return cfg


def inputs(cfg, B, seq_lens, seed, present=None):
    g = torch.Generator().manual_seed(seed)
    fc = cfg.feature_config
    names = fc['user_features'] + fc['item_features'] + fc['context_features']
    non_seq = {n: torch.randn(B, 1, generator=g, dtype=tf_shim.FLOAT) for n in names}
    seq = {n: torch.randn(B, L, 64, generator=g, dtype=tf_shim.FLOAT) for n, L in zip(fc['sequence_features'], seq_lens)
           if present is None or n in present}
    return non_seq, seq


def run_model_case(C, M, name, num_layers, pyramid, seq_lens, seed, present=None, drop_non_seq=False, ratios=None):
    cfg = small_config(C, num_layers, pyramid)
    if ratios is not None:
        cfg.pyramid_ratios = list(ratios)
    tf_shim.set_seed(seed)
    model = M.OneTransModel(cfg)
    non_seq, seq = inputs(cfg, 3, seq_lens, seed + 100, present)
    if drop_non_seq:      # no CONFIGURED feature present -> the zeros branch of OT/model.py:249-251 (an empty dict raises instead, see defects)
        non_seq = {'not_a_configured_feature': non_seq['price']}
    model(non_seq, seq, training=False)                       # builds every weight the run touches
    perturb(model, seed + 200)
    P = export_params(model, cfg)
    model.reset_kv_cache()
    tokens = model.tokenizer(non_seq, seq)
    out = model(non_seq, seq, training=False, use_kv_cache=False)
    arrays = {f'{name}/in/non_seq/{k}': v for k, v in non_seq.items()}
    arrays.update({f'{name}/in/seq/{k}': v for k, v in seq.items()})
    arrays.update({f'{name}/param/{k}': v for k, v in P.items()})
    arrays[f'{name}/out/tokens'] = tokens
    for t, v in out.items():
        arrays[f'{name}/out/prob/{t}'] = v
    paths = weight_paths(model)
    meta = {'weight_paths': paths, 'hidden_dim': cfg.hidden_dim, 'num_heads': cfg.num_heads, 'ffn_dim': cfg.ffn_dim, 'num_layers': cfg.num_layers,
            'num_ns_tokens': cfg.num_ns_tokens, 'pyramid_enabled': cfg.pyramid_enabled, 'pyramid_ratios': list(cfg.pyramid_ratios),
            'seq_lens': list(seq_lens), 'present': sorted(seq), 'total_len': int(tokens.shape[1]),
            'total_parameters': int(model.get_model_info()['total_parameters']),
            'kv_cache_len_after_call': int(model.kv_cache[0].shape[1])}
    return arrays, meta


def run_block_case(C, M, seed):
    cfg = small_config(C, 1, False)
    tf_shim.set_seed(seed)
    blk = M.OneTransBlock(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(2, 9, cfg.hidden_dim, generator=g, dtype=tf_shim.FLOAT)
    blk(x, training=False)
    perturb(blk, seed + 2)
    P = export_block(blk, 'blocks.0.')
    y, (k, v) = blk(x, training=False)
    arrays = {'block/in/x': x, 'block/out/y': y, 'block/out/k': k, 'block/out/v': v}
    arrays.update({f'block/param/{n}': t for n, t in P.items()})
    return arrays, {'seq_len': 9, 'num_ns_tokens': cfg.num_ns_tokens}


def weight_paths(root):
    """``(path, shape)`` of every built float weight of a reference layer tree in ATTRIBUTE-ASSIGNMENT order - the order Keras tracks
    sub-layers and therefore the order of ``model.weights`` / ``get_weights()`` / an ``.h5`` weight file (a layer's own variables come
    before its sub-layers; none of the reference's layers has both).  ``layers/i`` of a Sequential is spelled ``dense`` / ``dense_1``."""
    out = []

    def visit(obj, path):
        if isinstance(obj, torch.Tensor):
            if id(obj) in tf_shim._VARIABLES:
                out.append(('/'.join(path), list(obj.shape)))
        elif isinstance(obj, tf_shim.Sequential):
            dense_i = 0
            for layer in obj.layers:
                if isinstance(layer, tf_shim.Dense):
                    visit(layer, path + ['dense' if dense_i == 0 else f'dense_{dense_i}'])
                    dense_i += 1
        elif isinstance(obj, tf_shim.Layer):
            for k, v in vars(obj).items():
                if k not in ('_built', 'config'):
                    visit(v, path + [k])
        elif isinstance(obj, (list, tuple)):
            for i, v in enumerate(obj):
                visit(v, path + [str(i)])
        elif isinstance(obj, dict):
            for k, v in obj.items():
                visit(v, path + [str(k)])
    visit(root, [])
    return out


def run_gradient_case(C, M, name, seed):
    """Gradients of the summed BCE (OT/train.py:124-128; the loss expression is the oracle's restatement of Keras BinaryCrossentropy)
    through the REFERENCE's forward graph, by torch autograd on the shim's tensors - what ``tape.gradient`` (OT/train.py:131) returns."""
    from oracle import onetrans_oracle as O
    cfg = small_config(C, 2, False)
    tf_shim.set_seed(seed)
    model = M.OneTransModel(cfg)
    non_seq, seq = inputs(cfg, 3, (5, 4, 3), seed + 100)
    with torch.no_grad():
        model(non_seq, seq, training=False)
        perturb(model, seed + 200)
    live = [w for w in model.trainable_weights]
    for w in live:
        w.requires_grad_(True)
    g = torch.Generator().manual_seed(seed + 300)
    labels = {t: (torch.rand(3, 1, generator=g) < 0.5).to(tf_shim.FLOAT) for t in cfg.tasks}
    out = model(non_seq, seq, training=False)
    loss = O.bce_loss(out, labels, cfg.tasks)
    loss.backward()
    P = export_params(model, cfg)
    saved = {id(w): (w.grad if w.grad is not None else torch.zeros_like(w)) for w in live}
    for w in live:
        w.requires_grad_(False)
    swap = {id(w): w.detach().clone() for w in live}
    with torch.no_grad():
        for w in live:
            w.copy_(saved[id(w)])          # export the gradients with the parameter exporter, then put the weights back
        G = {k: v.detach().clone() for k, v in export_params(model, cfg).items()}
        for w in live:
            w.copy_(swap[id(w)])
    arrays = {f'{name}/in/non_seq/{k}': v for k, v in non_seq.items()}
    arrays.update({f'{name}/in/seq/{k}': v for k, v in seq.items()})
    arrays.update({f'{name}/in/label/{k}': v for k, v in labels.items()})
    arrays.update({f'{name}/param/{k}': v.detach() for k, v in P.items()})
    arrays.update({f'{name}/grad/{k}': v for k, v in G.items()})
    arrays[f'{name}/out/loss'] = loss.detach().reshape(1)
    return arrays, {'hidden_dim': cfg.hidden_dim, 'num_heads': cfg.num_heads, 'ffn_dim': cfg.ffn_dim, 'num_layers': cfg.num_layers,
                    'num_ns_tokens': cfg.num_ns_tokens, 'pyramid_enabled': cfg.pyramid_enabled, 'pyramid_ratios': list(cfg.pyramid_ratios)}


def expected_failure(fn):
    try:
        fn()
    except Exception as e:            # noqa: BLE001 - the point is to record what the reference raises
        return f'{type(e).__name__}: {e}'
    return None


def main():
    C, M = import_reference()
    facts = {'config': {}, 'scheduler': {}, 'cases': {}, 'defects': {}}
    for name in ('small', 'default', 'large'):
        cfg = C.get_model_config(name)
        facts['config'][name] = {k: v for k, v in cfg.to_dict().items() if isinstance(v, (int, float, str, bool, list, dict))}
    sched = M.PyramidScheduler(C.OneTransConfig())
    for L0 in (1, 2, 7, 21, 100, 272, 544, 1202, 2048):
        facts['scheduler'][str(L0)] = [sched.get_layer_config(l, L0).get('keep_len') for l in range(9)]
    lc = sched.get_layer_config(0, 21)
    facts['scheduler']['query_indices_layer0_len21'] = lc['query_indices']
    arrays = {}
    for name, kw in {'A_pyramid_off_2_blocks': dict(num_layers=2, pyramid=False, seq_lens=(5, 4, 3), seed=1),
                     'B_pyramid_on_1_block': dict(num_layers=1, pyramid=True, seq_lens=(6, 3, 4), seed=2),
                     'D_missing_sequence': dict(num_layers=1, pyramid=False, seq_lens=(4, 5, 3), seed=4, present=('click_seq', 'purchase_seq')),
                     'F_no_non_seq_features': dict(num_layers=1, pyramid=False, seq_lens=(3, 2, 4), seed=6, drop_non_seq=True),
                     'I_pyramid_keeps_one_token': dict(num_layers=1, pyramid=True, seq_lens=(2, 1, 1), seed=7, ratios=(0.05,)),
                     'J_only_last_sequence': dict(num_layers=1, pyramid=True, seq_lens=(2, 2, 6), seed=8, present=('purchase_seq',))}.items():
        a, meta = run_model_case(C, M, name, **kw)
        arrays.update(a)
        facts['cases'][name] = meta
    a, meta = run_block_case(C, M, 3)
    arrays.update(a)
    facts['cases']['block'] = meta
    # training=True: where the reference places its two dropouts per block (OT/model.py:193,198) - the shim's Dropout draws
    # torch.rand(shape, fp32) >= rate from a generator reseeded right before the call, the stream the oracle's _dropout consumes
    cfg = small_config(C, 2, False)
    tf_shim.set_seed(12)
    drop_model = M.OneTransModel(cfg)
    non_seq, seq = inputs(cfg, 3, (5, 4, 3), 112)
    with torch.no_grad():
        drop_model(non_seq, seq, training=False)
        perturb(drop_model, 212)
        tf_shim.set_seed(777)
        out = drop_model(non_seq, seq, training=True)
    arrays.update({f'L_dropout/in/non_seq/{k}': v for k, v in non_seq.items()})
    arrays.update({f'L_dropout/in/seq/{k}': v for k, v in seq.items()})
    arrays.update({f'L_dropout/param/{k}': v for k, v in export_params(drop_model, cfg).items()})
    arrays.update({f'L_dropout/out/prob/{t}': v for t, v in out.items()})
    facts['cases']['L_dropout'] = {'hidden_dim': cfg.hidden_dim, 'num_heads': cfg.num_heads, 'ffn_dim': cfg.ffn_dim, 'num_layers': cfg.num_layers,
                                   'num_ns_tokens': cfg.num_ns_tokens, 'pyramid_enabled': False, 'pyramid_ratios': list(cfg.pyramid_ratios),
                                   'dropout_rate': cfg.dropout_rate, 'dropout_seed': 777}
    a, meta = run_gradient_case(C, M, 'K_gradients', 11)
    arrays.update(a)
    facts['cases']['K_gradients'] = meta

    from tests.helpers import REFERENCE_KERNEL_CASES, REFERENCE_CPU_CASES
    for name, spec in {**REFERENCE_KERNEL_CASES, **REFERENCE_CPU_CASES}.items():
        a, meta = run_kernel_shape_case(C, M, name, spec)
        arrays.update(a)
        facts['cases'][name] = meta

    # parameter counts of the BASELINE configurations as the reference's own get_model_info reports them after one forward pass
    # (Keras builds weights lazily; pyramid off so that every block sees every position and builds every dedicated weight)
    facts['param_counts'] = {}
    for tag, preset, L_ns, lens in (('small_ns32_512', 'small', 32, (170, 170, 170)), ('small_ns16_256', 'small', 16, (86, 84, 84))):
        cfg = C.get_model_config(preset)
        cfg.num_ns_tokens, cfg.pyramid_enabled = L_ns, False
        tf_shim.set_seed(0)
        big = M.OneTransModel(cfg)
        non_seq, seq = inputs(cfg, 1, lens, 0)
        with torch.no_grad():
            big(non_seq, seq, training=False)
        facts['param_counts'][tag] = {'preset': preset, 'num_ns_tokens': L_ns, 'total_parameters': int(big.get_model_info()['total_parameters'])}
        del big

    # OT/data_loader.py:68-101, the reference's numpy code
    DL = importlib.import_module('ot_ref.data_loader')
    dcfg = C.OneTransConfig()
    dcfg.max_seq_len = 6
    sp = DL.SequenceProcessor(dcfg)
    g = torch.Generator().manual_seed(7)
    for tag, n in (('empty', 0), ('short', 4), ('exact', 6), ('long', 11)):
        raw = torch.randn(n, 64, generator=g, dtype=torch.float64).numpy()
        arrays[f'seqproc/{tag}/in'] = torch.from_numpy(raw)
        arrays[f'seqproc/{tag}/out'] = torch.from_numpy(np.asarray(sp.process_sequence(raw, 'click_seq'), dtype=np.float64))
    facts['cases']['seqproc'] = {'max_seq_len': 6}

    # OT/examples/inference_example.py cannot be imported (its annotations use Dict / Tuple without importing them, D11), but its method
    # bodies are plain Python: compile the reference's own function definitions out of the file and run them on a bare object
    import ast
    import typing
    tree = ast.parse(open(os.path.join(REF, 'examples', 'inference_example.py')).read())
    klass = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'OneTransInferenceEngine')
    wanted = [n for n in klass.body if isinstance(n, ast.FunctionDef) and n.name in ('_update_stats', 'get_stats', 'reset_stats', 'preprocess_input')]
    ns = {'Dict': typing.Dict, 'Tuple': typing.Tuple, 'List': typing.List, 'np': np}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), 'inference_example.py', 'exec'), ns)
    Eng = type('Eng', (), {k: ns[k] for k in ('_update_stats', 'get_stats', 'reset_stats', 'preprocess_input')})
    eng = Eng()
    eng.config = dcfg                                   # max_seq_len = 6
    eng.reset_stats()
    trace = []
    for success, latency, n in ((True, 12.0, 1), (True, 30.0, 4), (False, 0.0, 2), (True, 5.5, 3)):
        eng._update_stats(success, latency, n)
        trace.append(dict(eng.get_stats()))
    eng.reset_stats()
    rng2 = np.random.default_rng(5)
    seqs = {'click_seq': rng2.standard_normal((4, 64)), 'cart_seq': rng2.standard_normal((9, 64)).tolist(), 'purchase_seq': rng2.standard_normal((6, 64))}
    ns_out, seq_out = eng.preprocess_input({'user_id': 1.0}, {'item_id': 2.0, 'price': 3.5}, {'time': 0.25}, seqs)
    facts['cases']['inference_engine'] = {'stats_trace': trace, 'stats_after_reset': dict(eng.get_stats()), 'merged_non_seq': ns_out,
                                          'calls': [[True, 12.0, 1], [True, 30.0, 4], [False, 0.0, 2], [True, 5.5, 3]]}
    for k, v in seqs.items():
        arrays[f'engine/in/{k}'] = torch.from_numpy(np.asarray(v, dtype=np.float64))
        arrays[f'engine/out/{k}'] = torch.from_numpy(np.asarray(seq_out[k], dtype=np.float64))

    # OT/train.py:186-279 - the epoch loop (best-model / periodic / final checkpoints, early stopping, history, metric resets) is plain
    # Python around train_step / val_step / save_model: compile the reference's own `train` method and drive it with scripted losses
    ttree = ast.parse(open(os.path.join(REF, 'train.py')).read())
    tklass = next(n for n in ttree.body if isinstance(n, ast.ClassDef) and n.name == 'OneTransTrainer')
    train_def = next(n for n in tklass.body if isinstance(n, ast.FunctionDef) and n.name == 'train')
    import time as _time
    tns = {'Dict': typing.Dict, 'Optional': typing.Optional, 'DataLoader': object, 'np': np, 'time': _time, 'print': lambda *a, **k: None}
    exec(compile(ast.Module(body=[train_def], type_ignores=[]), 'train.py', 'exec'), tns)

    class _Num:
        def __init__(self, v): self.v = v
        def numpy(self): return self.v

    class _Metric:
        def __init__(self, log, name): self.log, self.name, self.n = log, name, 0
        def result(self): return _Num(float(self.n))
        def reset_states(self): self.log.append('reset:' + self.name); self.n += 1

    class _Loader:
        def __init__(self, n): self.n = n
        def get_train_dataset(self): return list(range(self.n))
        def get_val_dataset(self): return list(range(2))

    def drive(val_script, epochs, save_freq, patience, with_val=True):
        log = []
        T = type('T', (), {'train': tns['train']})
        t = T()
        t.config = 'cfg'
        t.history = {'train_loss': [], 'val_loss': [], 'train_metrics': {}, 'val_metrics': {}}
        t.train_metrics = {'ctr_auc': _Metric(log, 'train')}
        t.val_metrics = {'ctr_auc': _Metric(log, 'val')}
        state = {'epoch': -1, 'step': 0}
        def train_step(batch):
            if batch == 0:
                state['epoch'] += 1
            return {'total_loss': _Num(1.0 / (1 + state['epoch']) + 0.01 * batch)}
        def val_step(batch):
            return {'total_loss': _Num(val_script[state['epoch']] + 0.001 * batch)}
        t.train_step, t.val_step = train_step, val_step
        t.save_model = lambda name: log.append('save:' + name)
        hist = t.train(_Loader(3), _Loader(3) if with_val else None, epochs=epochs, save_freq=save_freq, early_stopping_patience=patience)
        return {'log': log, 'train_loss': [float(v) for v in hist['train_loss']], 'val_loss': [float(v) for v in hist['val_loss']],
                'epochs_run': len(hist['train_loss']), 'metric_epochs': sorted(int(k) for k in hist['val_metrics'])}
    facts['cases']['trainer_loop'] = {
        'early_stop': dict(args=dict(val_script=[1.0, 0.8, 0.9, 0.85, 0.95, 0.7, 0.6], epochs=7, save_freq=2, patience=3), **drive([1.0, 0.8, 0.9, 0.85, 0.95, 0.7, 0.6], 7, 2, 3)),
        'runs_out': dict(args=dict(val_script=[1.0, 0.9, 0.8, 0.7], epochs=4, save_freq=1, patience=5), **drive([1.0, 0.9, 0.8, 0.7], 4, 1, 5)),
        'no_val': dict(args=dict(val_script=[0, 0, 0], epochs=3, save_freq=3, patience=5, with_val=False), **drive([0, 0, 0], 3, 3, 5, with_val=False))}

    # OT/evaluate.py:131-169 (A/B arithmetic) and :231-282 (placeholder feature importance): plain Python around evaluate_offline
    etree = ast.parse(open(os.path.join(REF, 'evaluate.py')).read())
    eklass = next(n for n in etree.body if isinstance(n, ast.ClassDef) and n.name == 'OneTransEvaluator')
    edefs = [n for n in eklass.body if isinstance(n, ast.FunctionDef) and n.name in ('evaluate_ab_test', 'analyze_feature_importance')]
    ens = {'Dict': typing.Dict, 'Any': typing.Any, 'DataLoader': object, 'tf': tf_shim, 'print': lambda *a, **k: None}
    exec(compile(ast.Module(body=edefs, type_ignores=[]), 'evaluate.py', 'exec'), ens)
    EV = type('EV', (), {k: ens[k] for k in ('evaluate_ab_test', 'analyze_feature_importance')})
    ab_cases = {}
    for tag, (c, t) in {'gain': (0.70, 0.721), 'small': (0.70, 0.703), 'loss': (0.8, 0.75), 'zero_control': (0.0, 0.5)}.items():
        ev = EV()
        ev.evaluate_offline = lambda loader, kind, _v={'control': c, 'treatment': t}: {'ctr_auc': _v[loader]}
        r = ev.evaluate_ab_test('control', 'treatment')
        ab_cases[tag] = {'control': c, 'treatment': t, 'result': {k: (bool(v) if isinstance(v, (bool, np.bool_)) else v) for k, v in r.items()}}
    ev = EV()
    ev.config = C.OneTransConfig()
    ev.model = lambda *a, **k: {}
    ev.evaluate_offline = lambda loader, kind: {'ctr_auc': 0.7}
    class _DS:
        def get_test_dataset(self): return []
    facts['cases']['evaluator_logic'] = {'ab': ab_cases, 'feature_importance': ev.analyze_feature_importance(_DS())}

    # OT/data_loader.py:13-65 FeatureProcessor: pandas statistics, numpy standardisation, tf.one_hot
    import pandas as pd
    rng = np.random.default_rng(3)
    table = {'price': rng.uniform(0, 1000, 50).tolist(), 'age': rng.integers(18, 70, 50).astype(float).tolist(),
             'user_id': rng.integers(0, 40, 50).tolist(), 'category': rng.integers(0, 7, 50).tolist()}
    fp = DL.FeatureProcessor(dcfg)
    fp.fit(pd.DataFrame(table))
    probe = {'price': [0.0, 250.0, 999.0, 5000.0, -4000.0], 'age': [18.0, 44.0, 69.0], 'ctr': [0.1, 0.2]}
    facts['cases']['featproc'] = {
        'table': table, 'feature_stats': {k: {kk: float(vv) for kk, vv in v.items()} for k, v in fp.feature_stats.items()},
        'vocab_sizes': {k: int(v) for k, v in fp.vocab_sizes.items()}, 'probe': probe,
        'numerical': {k: np.asarray(fp.process_numerical_feature(k, np.array(v))).tolist() for k, v in probe.items()},
        'one_hot_category': np.asarray(fp.process_categorical_feature('category', np.array([0, 3, 6]))).tolist(),
        'unknown_categorical_passthrough': np.asarray(fp.process_categorical_feature('brand', np.array([5, 9]))).tolist()}

    # the defects of SURVEY.md §A.3 that the shimmed reference reproduces
    def d2():
        cfg = small_config(C, 2, True)
        non_seq, seq = inputs(cfg, 2, (6, 3, 4), 9)
        M.OneTransModel(cfg)(non_seq, seq, training=False)
    facts['defects']['D2_pyramid_on_two_blocks'] = expected_failure(d2)

    def d9():
        cfg = small_config(C, 1, False)
        non_seq, seq = inputs(cfg, 2, (6, 3, 4), 9)
        non_seq['user_id'] = torch.randint(0, 100, (2, 1))
        M.OneTransModel(cfg)(non_seq, seq, training=False)
    facts['defects']['D9_integer_ids_in_concat'] = expected_failure(d9)

    def d6():      # second call with use_kv_cache=True: cached keys in front, mask still [L, L] (OT/model.py:95-98, 109-110)
        cfg = small_config(C, 1, False)
        non_seq, seq = inputs(cfg, 2, (6, 3, 4), 9)
        m = M.OneTransModel(cfg)
        m(non_seq, seq, training=False, use_kv_cache=True)
        m(non_seq, seq, training=False, use_kv_cache=True)
    facts['defects']['D6_kv_cache_second_call'] = expected_failure(d6)

    def no_seq():  # OT/model.py:274-275: the empty-sequence fallback indexes list(features.values())[0] of an EMPTY dict
        cfg = small_config(C, 1, False)
        non_seq, _ = inputs(cfg, 2, (6, 3, 4), 9)
        M.OneTransModel(cfg)(non_seq, {}, training=False)
    facts['defects']['no_sequences_at_all'] = expected_failure(no_seq)

    def no_ns():   # OT/model.py:249-251 likewise for an empty non-sequence dict
        cfg = small_config(C, 1, False)
        _, seq = inputs(cfg, 2, (6, 3, 4), 9)
        M.OneTransModel(cfg)({}, seq, training=False)
    facts['defects']['no_non_seq_features_at_all'] = expected_failure(no_ns)

    def d7():      # how OT/train.py:118,163, OT/evaluate.py:87 and the example scripts call the model: ONE tuple argument
        cfg = small_config(C, 1, False)
        non_seq, seq = inputs(cfg, 2, (6, 3, 4), 9)
        M.OneTransModel(cfg)((non_seq, seq), training=True)
    facts['defects']['D7_tuple_call'] = expected_failure(d7)
    facts['defects']['D8_config_has_gradient_clip'] = hasattr(C.OneTransConfig(), 'gradient_clip')      # read by OT/train.py:134
    facts['defects']['D8_config_has_gradient_clip_norm'] = hasattr(C.OneTransConfig(), 'gradient_clip_norm')

    np.savez_compressed(os.path.join(HERE, 'reference_golden.npz'), **{k: v.detach().numpy() for k, v in arrays.items()})
    with open(os.path.join(HERE, 'reference_golden.json'), 'w') as f:
        json.dump(facts, f, indent=1, sort_keys=True)
    print('wrote reference_golden.npz', os.path.getsize(os.path.join(HERE, 'reference_golden.npz')), 'bytes;', len(arrays), 'arrays')
    print('defects reproduced:', json.dumps(facts['defects'], indent=1))


if __name__ == '__main__':
    main()
