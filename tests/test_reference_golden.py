"""The oracle against vectors produced by the REFERENCE'S OWN CODE: tests/golden/reference_golden.{npz,json} come from
/root/reference/rank/scaling_up/oneTrans/practice/{config,model}.py, imported unmodified and executed over a shim of the TensorFlow
ops they call (oracle/tf_shim.py; generator tests/golden/make_reference_golden.py).  The control flow that decides results - weight
selection per position, concat orders, mask, [SEP] placement, pyramid indices and gathers, last-token heads - is the reference's.

The reference applies dedicated weights to positions < num_ns_tokens (OT/model.py:69-74, 155-157): the oracle's literal mode
``ns_param_alignment='head_literal'`` + ``query_mode='literal_gather'`` is the one compared; the repaired modes the product uses are
tied to it by T4 / T5 in tests/test_oracle.py (tail-only == compute-all-then-gather, grouped == per-token loop)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import onetrans_oracle as O
import recommend_b200 as R

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
Z = np.load(os.path.join(HERE, 'reference_golden.npz'))
FACTS = json.load(open(os.path.join(HERE, 'reference_golden.json')))


def _group(prefix):
    return {k[len(prefix):]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith(prefix)}


def _oracle_cfg(meta, literal=True):
    cfg = O.OracleConfig(hidden_dim=meta['hidden_dim'], num_layers=meta['num_layers'], num_heads=meta['num_heads'], ffn_dim=meta['ffn_dim'],
                         num_ns_tokens=meta['num_ns_tokens'], pyramid_enabled=meta['pyramid_enabled'], pyramid_ratios=meta['pyramid_ratios'],
                         ns_param_alignment='head_literal', dropout_rate=0.1)
    return cfg


@pytest.mark.parametrize('case', ['A_pyramid_off_2_blocks', 'B_pyramid_on_1_block', 'D_missing_sequence', 'F_no_non_seq_features',
                                  'I_pyramid_keeps_one_token', 'J_only_last_sequence'])
def test_oracle_equals_the_reference_model_call(case):
    meta = FACTS['cases'][case]
    cfg = _oracle_cfg(meta)
    P, non_seq, seq = _group(f'{case}/param/'), _group(f'{case}/in/non_seq/'), _group(f'{case}/in/seq/')
    assert sorted(seq) == meta['present']
    P.setdefault('tokenizer.sep_embedding', torch.zeros(1, meta['hidden_dim'], dtype=torch.float64))   # unbuilt in the reference when no [SEP] is placed
    tokens = O.tokenizer_forward(P, cfg, non_seq, seq)
    want_tokens = torch.from_numpy(Z[f'{case}/out/tokens'])
    assert tokens.shape == want_tokens.shape and tokens.shape[1] == meta['total_len']
    assert torch.allclose(tokens, want_tokens, rtol=0, atol=1e-13)                        # layout + projections: S first, [SEP]s, NS last
    for mode, loop in (('literal_gather', True), ('literal_gather', False), ('tail_only', False)):
        out = O.model_forward(P, cfg, non_seq, seq, training=False, query_mode=mode, literal_loop=loop)
        for t in cfg.tasks:
            want = torch.from_numpy(Z[f'{case}/out/prob/{t}'])
            assert out[t].shape == want.shape == (3, 1)
            assert torch.allclose(out[t], want, rtol=0, atol=1e-12), (case, mode, loop, float((out[t] - want).abs().max()))
    # every parameter the run touched, counted the way get_model_info does (OT/model.py:399-408)
    touched = sum(int(np.count_nonzero(Z[k]) > 0) * Z[k].size for k in Z.files if k.startswith(f'{case}/param/'))
    assert touched <= meta['total_parameters']


def test_oracle_equals_the_reference_block_and_its_kv():
    meta = FACTS['cases']['block']
    cfg = O.OracleConfig(hidden_dim=16, num_layers=1, num_heads=4, ffn_dim=24, num_ns_tokens=meta['num_ns_tokens'], ns_param_alignment='head_literal')
    P = _group('block/param/')
    x = torch.from_numpy(Z['block/in/x'])
    y = O.block_forward(P, 0, cfg, x, keep=meta['seq_len'], query_mode='literal_gather', literal_loop=True)
    assert torch.allclose(y, torch.from_numpy(Z['block/out/y']), rtol=0, atol=1e-13)
    xn = O.rmsnorm(x, P['blocks.0.norm1.scale'])
    _, (k, v) = O.mixed_mha(P, 'blocks.0.attention.', cfg, xn, meta['seq_len'], return_kv=True)
    B, L = x.shape[:2]
    assert torch.allclose(k.reshape(B, L, -1), torch.from_numpy(Z['block/out/k']).reshape(B, L, -1), rtol=0, atol=1e-13)   # (k, v) as MixedMHA returns them
    assert torch.allclose(v.reshape(B, L, -1), torch.from_numpy(Z['block/out/v']).reshape(B, L, -1), rtol=0, atol=1e-13)


def test_scheduler_and_config_facts_of_the_reference():
    ratios = FACTS['config']['default']['pyramid_ratios']
    for L0, keeps in FACTS['scheduler'].items():
        if not L0.isdigit():
            continue
        for l, k in enumerate(keeps):                                # None past the ratio list (OT/model.py:289-290)
            assert O.reference_keep_len(l, int(L0), ratios) == k
            got = R.PyramidScheduler(R.OneTransConfig()).get_layer_config(l, int(L0))
            assert got.get('keep_len') == k
    assert FACTS['scheduler']['query_indices_layer0_len21'] == O.reference_query_indices(0, 21, ratios) == list(range(11, 21))
    for name, ref in FACTS['config'].items():                        # the attribute bag of OT/config.py, value by value
        mine = R.get_model_config(name).to_dict()
        for key, value in ref.items():
            assert key in mine and mine[key] == value, (name, key, mine.get(key), value)


def test_defects_the_reference_shows_when_executed():
    """SURVEY.md §A.3 D2 and D9, observed by running the reference (the repairs the oracle and the product adopt start here)."""
    assert 'is not in [0,' in FACTS['defects']['D2_pyramid_on_two_blocks']          # layer 1 gathers original-length indices
    assert 'different dtypes' in FACTS['defects']['D9_integer_ids_in_concat']       # int ids next to float features
    assert 'ranks of all input tensors should match' in FACTS['defects']['D6_kv_cache_second_call']   # the cache holds the 4-D reshaped k / v
    assert FACTS['defects']['no_sequences_at_all'].startswith('IndexError')          # OT/model.py:274-275 indexes an empty dict's values
    assert FACTS['defects']['no_non_seq_features_at_all'].startswith('IndexError')   # OT/model.py:249-251 likewise
    assert FACTS['defects']['D7_tuple_call'].startswith('TypeError')                 # model((non_seq, seq)) as train.py / evaluate.py call it
    assert FACTS['defects']['D8_config_has_gradient_clip'] is False and FACTS['defects']['D8_config_has_gradient_clip_norm'] is True
    assert R.OneTransConfig().gradient_clip_norm == 90.0 and not hasattr(R.OneTransConfig(), 'gradient_clip')
    # after a call the model holds the LAST block's (k, v) as its "cache" (OT/model.py:366-368) - per-layer caching needs D6
    assert FACTS['cases']['A_pyramid_off_2_blocks']['kv_cache_len_after_call'] == FACTS['cases']['A_pyramid_off_2_blocks']['total_len']


def test_sequence_processor_equals_the_reference_numpy_code():
    """OT/data_loader.py:68-101 executed (numpy): newest ``max_seq_len`` events kept, shorter sequences padded IN FRONT, an empty one
    becomes all zeros - against ``recommend_b200.data.SequenceProcessor`` and the serving wrapper's ``preprocess_input``."""
    cfg = R.OneTransConfig()
    cfg.max_seq_len = FACTS['cases']['seqproc']['max_seq_len']
    sp = R.SequenceProcessor(cfg)
    eng = R.OneTransInferenceEngine.__new__(R.OneTransInferenceEngine)       # preprocess_input only needs the config
    eng.config, eng.pad_sequences = cfg, True
    for tag in ('empty', 'short', 'exact', 'long'):
        raw, want = torch.from_numpy(Z[f'seqproc/{tag}/in']), torch.from_numpy(Z[f'seqproc/{tag}/out'])
        got = sp.process_sequence(raw)
        assert got.shape == want.shape == (cfg.max_seq_len, 64) and torch.equal(got.double(), want.float().double()), tag
        if raw.shape[0]:
            _, seq = eng.preprocess_input({}, {}, {}, {'click_seq': raw})
            assert torch.equal(seq['click_seq'].double(), want.float().double()), tag


@pytest.mark.parametrize('case', ['G_d256_pyramid_on_1_block', 'H_d256_pyramid_off_2_blocks'])
def test_kernel_shape_cases_rebuild_from_seeds_and_match_the_reference(case):
    """The CPU twin of tests/test_gpu_model.py::test_product_equals_the_reference_outputs: same rebuilt weights / inputs, the oracle
    in place of the CUDA model, against the probabilities the reference's own code produced (fp64: 1e-12)."""
    from tests.helpers import reference_case_inputs, reference_case_checksum
    spec = FACTS['cases'][case]
    ocfg, P, non_seq, seq = reference_case_inputs(spec)
    assert reference_case_checksum(P, non_seq, seq) == pytest.approx(spec['checksum'], rel=1e-13)     # the seeds rebuild what the reference saw
    f64 = lambda d: {k: v.double() for k, v in d.items()}
    out = O.model_forward(f64(P), ocfg, f64(non_seq), f64(seq), query_mode='literal_gather')
    for t in ocfg.tasks:
        want = torch.from_numpy(Z[f'{case}/out/prob/{t}'])
        assert torch.allclose(out[t], want, rtol=0, atol=1e-12), float((out[t] - want).abs().max())


def test_feature_processor_equals_the_reference_pandas_numpy_code():
    """OT/data_loader.py:13-65 executed (pandas statistics, numpy z-score + clip, one-hot) against ``recommend_b200.FeatureProcessor``."""
    f = FACTS['cases']['featproc']
    fp = R.FeatureProcessor(R.OneTransConfig())
    fp.fit(f['table'])
    assert fp.vocab_sizes == f['vocab_sizes'] and set(fp.feature_stats) == set(f['feature_stats'])
    for name, st in f['feature_stats'].items():
        for key, value in st.items():
            assert fp.feature_stats[name][key] == pytest.approx(value, rel=1e-12), (name, key)
    for name, values in f['probe'].items():
        got = fp.process_numerical_feature(name, values)
        assert torch.allclose(got, torch.tensor(f['numerical'][name], dtype=torch.float64), rtol=1e-12, atol=1e-12), name
    assert fp.process_categorical_feature('category', [0, 3, 6]).tolist() == f['one_hot_category']
    assert fp.process_categorical_feature('brand', [5, 9]).tolist() == f['unknown_categorical_passthrough']     # unfitted feature: unchanged


def test_keras_weight_order_follows_the_reference_object_graph():
    """``state.keras_weight_list`` (what ``OneTransTrainer.save_model`` writes and ``load_keras_weight_list`` reads) lists the weights in
    the attribute-assignment order of the reference's constructors - recorded from the LIVE reference model (case A), names and shapes."""
    from recommend_b200 import state
    meta = FACTS['cases']['A_pyramid_off_2_blocks']
    cfg = R.OneTransConfig()
    cfg.hidden_dim, cfg.num_heads, cfg.ffn_dim, cfg.num_layers, cfg.num_ns_tokens = (meta['hidden_dim'], meta['num_heads'], meta['ffn_dim'],
                                                                                     meta['num_layers'], meta['num_ns_tokens'])
    mine = [(name, list(a.shape)) for name, a in state.keras_weight_list(R.OneTransModel(cfg))]
    want = [(name, shape) for name, shape in meta['weight_paths']]
    assert len(mine) == len(want) == 80
    assert mine == want
    assert sum(int(np.prod(s)) for _, s in want) == meta['total_parameters']


def test_parameter_counts_of_the_baseline_configurations():
    """The reference's own ``get_model_info`` after one forward pass at the BASELINE shapes (SURVEY.md §8d "Params / allreduce payload":
    OneTrans-S with 32 NS tokens = 143.6 M) against the product model's count."""
    assert FACTS['param_counts']['small_ns32_512']['total_parameters'] == 143601922
    for tag, ref in FACTS['param_counts'].items():
        cfg = R.get_model_config(ref['preset'])
        cfg.num_ns_tokens = ref['num_ns_tokens']
        with torch.device('meta'):
            model = R.OneTransModel(cfg)
        assert model.get_model_info()['total_parameters'] == ref['total_parameters'], tag


def test_baseline_config_1_forward_and_loss_match_the_reference_at_full_size():
    """BASELINE config 1 - OneTrans-S, CPU fp32, batch 32, 6 blocks, 256 sequence + 16 NS tokens - run by the reference's own code
    (pyramid off: it cannot prune past one block, D2).  The oracle in fp32 (the arithmetic type config 1 names) against the reference's
    fp64 probabilities, and the summed BCE of OT/train.py:84-87, 124-128 from both."""
    from tests.helpers import reference_case_inputs, reference_case_checksum
    case = 'C1_small_ns16_pyramid_off'
    spec = FACTS['cases'][case]
    ocfg, P, non_seq, seq = reference_case_inputs(spec)
    assert reference_case_checksum(P, non_seq, seq) == pytest.approx(spec['checksum'], rel=1e-12)
    assert spec['total_len'] == 272
    out = O.model_forward(P, ocfg, non_seq, seq)                               # fp32, vectorised, tail-only == everything (pyramid off)
    g = torch.Generator().manual_seed(5)
    for t in ocfg.tasks:
        want = torch.from_numpy(Z[f'{case}/out/prob/{t}'])
        assert out[t].dtype == torch.float32 and out[t].shape == want.shape == (32, 1)
        assert float((out[t].double() - want).abs().max()) < 5e-6
        y = (torch.rand(32, 1, generator=g) < 0.5).float()
        assert float(O.bce_loss({t: out[t]}, {t: y}, [t])) == pytest.approx(float(O.bce_loss({t: want}, {t: y.double()}, [t])), rel=1e-5)


def test_oracle_gradients_equal_autograd_through_the_reference_forward():
    """``tape.gradient`` of OT/train.py:131: the summed BCE differentiated through the reference's own forward graph (torch autograd on
    the shim's tensors) against the oracle's ``loss_and_grads`` in its literal mode - every parameter tensor, packed as the GPU parity
    tests compare them."""
    meta = FACTS['cases']['K_gradients']
    cfg = _oracle_cfg(meta)
    cfg.dropout_rate = 0.0
    P, G = _group('K_gradients/param/'), _group('K_gradients/grad/')
    non_seq, seq, labels = _group('K_gradients/in/non_seq/'), _group('K_gradients/in/seq/'), _group('K_gradients/in/label/')
    loss, grads, _ = O.loss_and_grads(P, cfg, non_seq, seq, labels, query_mode='literal_gather', literal_loop=True)
    assert float(loss) == pytest.approx(float(Z['K_gradients/out/loss'][0]), rel=1e-13)
    assert set(G) == set(P) and len(G) > 20
    for name, want in G.items():
        got = grads[name]
        assert got.shape == want.shape, name
        assert torch.allclose(got, want, rtol=0, atol=1e-12 + 1e-10 * float(want.abs().max())), (name, float((got - want).abs().max()))
    assert float(G['blocks.0.attention.Wq'][0].abs().max()) > 0 and float(G['tokenizer.sep_embedding'].abs().max()) > 0


def test_dropout_placement_equals_the_reference_in_training_mode():
    """training=True through the reference (OT/model.py:193,198: one Dropout layer applied to the attention branch, then to the FFN
    branch, per block, on the full-length tensors) with the mask stream seeded identically on both sides."""
    meta = FACTS['cases']['L_dropout']
    cfg = _oracle_cfg(meta)
    cfg.dropout_rate = meta['dropout_rate']
    P, non_seq, seq = _group('L_dropout/param/'), _group('L_dropout/in/non_seq/'), _group('L_dropout/in/seq/')
    out = O.model_forward(P, cfg, non_seq, seq, training=True, query_mode='literal_gather', literal_loop=True,
                          gen=torch.Generator().manual_seed(meta['dropout_seed']))
    plain = O.model_forward(P, cfg, non_seq, seq, training=False, query_mode='literal_gather')
    for t in cfg.tasks:
        want = torch.from_numpy(Z[f'L_dropout/out/prob/{t}'])
        assert torch.allclose(out[t], want, rtol=0, atol=1e-12), float((out[t] - want).abs().max())
        assert float((plain[t] - want).abs().max()) > 1e-4                      # the masks did something


def test_inference_engine_bookkeeping_equals_the_reference_method_bodies():
    """``_update_stats`` / ``get_stats`` / ``reset_stats`` / ``preprocess_input`` of OT/examples/inference_example.py:62-92, 186-219,
    compiled from the reference file (the module itself cannot be imported, D11) and run; against ``OneTransInferenceEngine``."""
    f = FACTS['cases']['inference_engine']
    cfg = R.OneTransConfig()
    cfg.max_seq_len = FACTS['cases']['seqproc']['max_seq_len']
    eng = R.OneTransInferenceEngine.__new__(R.OneTransInferenceEngine)
    eng.config, eng.pad_sequences = cfg, True
    eng.reset_stats()
    for (success, latency, n), want in zip(f['calls'], f['stats_trace']):
        eng._update_stats(success, latency, n)
        got = eng.get_stats()
        assert set(got) == set(want)
        for k in want:
            assert got[k] == pytest.approx(want[k], rel=1e-12), (k, got, want)
    eng.reset_stats()
    assert eng.get_stats() == f['stats_after_reset']
    seqs = {k[len('engine/in/'):]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith('engine/in/')}
    non_seq, seq = eng.preprocess_input({'user_id': 1.0}, {'item_id': 2.0, 'price': 3.5}, {'time': 0.25}, seqs)
    assert non_seq == f['merged_non_seq']
    for k, v in seq.items():
        assert torch.equal(v.double(), torch.from_numpy(Z[f'engine/out/{k}']).float().double()), k


@pytest.mark.parametrize('script', ['early_stop', 'runs_out', 'no_val'])
def test_trainer_epoch_loop_equals_the_reference_train_method(script, capsys):
    """OT/train.py:186-279 compiled from the reference file and driven with scripted losses: which checkpoints are written when
    (``best_model`` on improvement, ``model_epoch_k`` every ``save_freq``, ``final_model``), when early stopping ends the run, when the
    metrics are reset, and the loss history - against ``OneTransTrainer.train`` driven by the same script (no GPU involved)."""
    from recommend_b200.train import OneTransTrainer
    want = FACTS['cases']['trainer_loop'][script]
    a = want['args']
    log = []

    class Metrics:
        def __init__(self, name): self.name = name
        def result(self): return {}
        def reset_states(self): log.append('reset:' + self.name)

    class Loader:
        def get_train_dataset(self): return list(range(3))
        def get_val_dataset(self): return list(range(2))

    class Model:
        def train(self): pass
        def eval(self): pass

    t = OneTransTrainer.__new__(OneTransTrainer)
    t.config, t.model = None, Model()
    t.history = {'train_loss': [], 'val_loss': [], 'train_metrics': {}, 'val_metrics': {}}
    t.train_metrics, t.val_metrics = Metrics('train'), Metrics('val')
    state = {'epoch': -1}

    def train_step(batch):
        if batch == 0:
            state['epoch'] += 1
        return {'total_loss': torch.tensor(1.0 / (1 + state['epoch']) + 0.01 * batch, dtype=torch.float64)}
    t.train_step = train_step
    t.val_step = lambda batch: {'total_loss': torch.tensor(a['val_script'][state['epoch']] + 0.001 * batch, dtype=torch.float64)}
    t.save_model = lambda name: log.append('save:' + name)
    hist = t.train(Loader(), Loader() if a.get('with_val', True) else None, epochs=a['epochs'], save_freq=a['save_freq'],
                   early_stopping_patience=a['patience'], log_every=0)
    assert log == want['log']
    assert len(hist['train_loss']) == want['epochs_run']
    assert hist['train_loss'] == pytest.approx(want['train_loss'], rel=1e-12) and hist['val_loss'] == pytest.approx(want['val_loss'], rel=1e-12)
    assert sorted(hist['val_metrics']) == want['metric_epochs']


def test_evaluator_ab_arithmetic_and_placeholder_importance_equal_the_reference_methods():
    """OT/evaluate.py:131-169 and :231-282 compiled from the reference file and driven with stubbed ``evaluate_offline`` results."""
    from recommend_b200.evaluate import OneTransEvaluator
    f = FACTS['cases']['evaluator_logic']
    for tag, case in f['ab'].items():
        ev = OneTransEvaluator.__new__(OneTransEvaluator)
        ev.evaluate_offline = lambda loader, kind, _v={'control': case['control'], 'treatment': case['treatment']}: {'ctr_auc': _v[loader]}
        got = ev.evaluate_ab_test('control', 'treatment')
        assert set(got) == set(case['result']), tag
        for k, want in case['result'].items():
            assert (got[k] == pytest.approx(want, rel=1e-12)) if isinstance(want, float) else (got[k] == want), (tag, k, got[k], want)
    ev = OneTransEvaluator.__new__(OneTransEvaluator)
    ev.config = R.OneTransConfig()
    got = ev.analyze_feature_importance(None)
    assert set(got) == set(f['feature_importance'])
    for k, want in f['feature_importance'].items():
        assert got[k] == pytest.approx(want, rel=1e-12), k
