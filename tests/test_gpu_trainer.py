"""OneTransTrainer / OneTransEvaluator end to end on a small model (surface of OT/train.py:19-338 and OT/evaluate.py:22-416): the
loss falls and the AUC rises on data whose labels depend on the features, checkpoints reload bit-exactly, the evaluator's streaming
metrics equal the oracle's on the predictions it saw."""
import json

import numpy as np
import pytest
import torch

import recommend_b200 as R
from oracle import metrics_oracle as M

pytestmark = pytest.mark.gpu


def _config():
    cfg = R.get_model_config('small')
    cfg.num_layers, cfg.num_ns_tokens, cfg.max_seq_len, cfg.pyramid_schedule = 2, 4, 6, 'linear_to_ns'
    cfg.batch_size, cfg.dropout_rate = 64, 0.0
    cfg.optimizer_config['momentum'] = 0.9
    cfg.optimizer_config['dense_lr'] = 0.003
    return cfg


def _loader(cfg, n=(1024, 256, 256)):
    dl = R.DataLoader(cfg)
    dl.train_dataset = dl.create_sample_data(n[0], seed=1, label_signal=3.0)
    dl.val_dataset = dl.create_sample_data(n[1], seed=2, label_signal=3.0)
    dl.test_dataset = dl.create_sample_data(n[2], seed=3, label_signal=3.0)
    return dl


def test_trainer_learns_checkpoints_and_evaluates(tmp_path):
    torch.manual_seed(0)
    cfg = _config()
    dl = _loader(cfg)
    trainer = R.OneTransTrainer(cfg, str(tmp_path / 'models'))
    hist = trainer.train(dl, dl, epochs=4, save_freq=2, early_stopping_patience=5, log_every=0)
    print('history', hist['train_loss'], hist['val_loss'], [hist['val_metrics'][e]['ctr_auc'] for e in range(4)])
    assert len(hist['train_loss']) == 4 and hist['train_loss'][-1] < hist['train_loss'][0]
    assert max(hist['val_metrics'][e]['ctr_auc'] for e in range(4)) > 0.55 and set(hist['val_metrics'][0]) >= {'ctr_auc', 'cvr_recall', 'ctr_logloss'}
    for name in ('best_model', 'model_epoch_2', 'model_epoch_4', 'final_model'):
        assert (tmp_path / 'models' / name / 'model_weights.npz').exists() and (tmp_path / 'models' / name / 'config.json').exists()
    saved = json.load(open(tmp_path / 'models' / 'final_model' / 'training_history.json'))
    assert saved['train_loss'] == hist['train_loss']

    # reload: same predictions bit for bit (fp32 masters round-trip exactly through the .npz)
    batch = next(iter(dl.get_test_dataset()))
    dev = lambda part: {k: v.cuda() for k, v in part.items()}
    with torch.no_grad():
        want = trainer.model.eval()(dev(batch[0]), dev(batch[1]), training=False)
    model2, cfg2 = R.load_model_for_evaluation(str(tmp_path / 'models' / 'final_model'))
    with torch.no_grad():
        got = model2.eval()(dev(batch[0]), dev(batch[1]), training=False)
    assert all(torch.equal(want[t], got[t]) for t in cfg.tasks)
    t2 = R.OneTransTrainer(cfg, str(tmp_path / 'models2'))
    t2.load_model(tmp_path / 'models' / 'final_model')
    assert t2.history['val_loss'] == hist['val_loss']
    with torch.no_grad():
        got2 = t2.model.eval()(dev(batch[0]), dev(batch[1]), training=False)
    assert all(torch.equal(want[t], got2[t]) for t in cfg.tasks)

    # evaluator: streaming metrics == oracle on the predictions of the same pass
    ev = R.OneTransEvaluator(model2, cfg2)
    res = ev.evaluate_offline(dl, 'test')
    ys, ps = [], []
    with torch.no_grad():
        for b in dl.get_test_dataset():
            ps.append(model2(dev(b[0]), dev(b[1]), training=False)['ctr'].float().cpu().numpy().reshape(-1))
            ys.append(b[2]['ctr'].numpy().reshape(-1))
    y, p = np.concatenate(ys), np.concatenate(ps)
    assert res['ctr_auc'] == pytest.approx(M.keras_auc(y, p), abs=2e-6)
    assert res['ctr_auc_exact'] == pytest.approx(M.exact_auc(y, p), abs=1e-12)
    assert res['ctr_accuracy'] == pytest.approx(M.binary_accuracy(y, p), abs=1e-12)
    assert res['ctr_logloss'] == pytest.approx(M.binary_crossentropy(y, p), rel=1e-5)
    perf = res['performance']
    assert perf['total_samples'] == 256 and perf['throughput_samples_per_second'] > 0
    ab = ev.evaluate_ab_test(dl, dl)
    assert ab['absolute_improvement'] == 0 and ab['is_statistically_significant'] is False and ab['metric_name'] == 'ctr_auc'
    bench = ev.benchmark_performance(dl, num_batches=3, warmup_batches=1)
    assert bench['avg_inference_time_ms'] > 0 and bench['p99_inference_time_ms'] >= bench['p95_inference_time_ms'] and bench['total_batches_tested'] == 3
    imp = ev.analyze_feature_importance(dl)
    assert imp['click_seq'] == pytest.approx(0.2 / 1.7) and imp['age'] == pytest.approx(0.1 / 1.7)       # the reference's placeholder scores
    imp2 = ev.analyze_feature_importance(dl, method='ablation')
    assert max(imp2, key=imp2.get) in ('user_id', 'click_seq')                                            # the two features the labels depend on
    report = R.evaluate_model(str(tmp_path / 'models' / 'final_model'), dl, str(tmp_path / 'reports'))
    rep = json.load(open(report['report_path']))
    assert set(rep) == {'model_config', 'offline_evaluation', 'performance_benchmark', 'feature_importance', 'evaluation_timestamp', 'data_info'}
    assert rep['data_info'] == {'train_samples': 1024, 'val_samples': 256, 'test_samples': 256}
    # the serving wrapper reads the trainer's directory too
    eng = R.OneTransInferenceEngine(tmp_path / 'models' / 'final_model')
    assert eng.config.num_layers == 2


def test_trainer_rejects_adam_and_missing_files(tmp_path):
    cfg = _config()
    cfg.optimizer_config['dense_optimizer'] = 'adam'
    with pytest.raises(NotImplementedError):
        R.OneTransTrainer(cfg, str(tmp_path))
    with pytest.raises(FileNotFoundError):
        R.load_model_for_evaluation(str(tmp_path / 'nothing'))
