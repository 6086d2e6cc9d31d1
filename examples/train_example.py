"""Training / evaluation / inference walk-through on the sm_100a path - the counterpart of the reference's
OT/examples/train_example.py (basic_training_example :17-43, full_training_pipeline :46-72, evaluation_example :75-103,
model_inference_example :106-129).  Needs a B200: there is no CPU fallback.

    python examples/train_example.py [--epochs 3] [--samples 2048] [--model-dir ./example_models]

Shapes differ from the reference script where its own would not run here: it asks for hidden_dim 128 with 4 heads (head_dim 32);
the attention kernels are built for head_dim 64 and 96, so the small example model is d = 256, H = 4."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import recommend_b200 as R


def example_config():
    config = R.get_model_config('small')
    config.num_layers, config.num_ns_tokens, config.max_seq_len = 4, 8, 50          # OT/examples/train_example.py:22-27
    config.batch_size, config.pyramid_schedule = 256, 'linear_to_ns'
    config.optimizer_config['momentum'], config.optimizer_config['dense_lr'] = 0.9, 0.002
    return config


def basic_training_example(config):
    """One forward pass on a sample batch (OT/examples/train_example.py:17-43)."""
    model = R.OneTransModel(config).cuda()
    non_seq, seq, labels = R.create_sample_batch(config, 2, (10, 5, 7), ns_mode='ids')
    dev = lambda d: {k: v.cuda() for k, v in d.items()}
    predictions = model(dev(non_seq), dev(seq), training=True)
    print('forward ok:', {k: tuple(v.shape) for k, v in predictions.items()}, model.get_model_info())
    return model


def full_training_pipeline(config, epochs, samples, model_dir):
    """Loader -> trainer -> train (OT/examples/train_example.py:46-72)."""
    data_loader = R.DataLoader(config)
    data_loader.train_dataset = data_loader.create_sample_data(samples, seed=1, label_signal=3.0)
    data_loader.val_dataset = data_loader.create_sample_data(samples // 4, seed=2, label_signal=3.0)
    data_loader.test_dataset = data_loader.create_sample_data(samples // 4, seed=3, label_signal=3.0)
    trainer = R.OneTransTrainer(config, model_dir=model_dir)
    history = trainer.train(train_loader=data_loader, val_loader=data_loader, epochs=epochs, save_freq=1, early_stopping_patience=3)
    print('train loss per epoch:', [round(v, 4) for v in history['train_loss']])
    return trainer, data_loader


def evaluation_example(trainer, data_loader, model_dir):
    """Offline metrics, forward benchmark and the JSON report (OT/examples/train_example.py:75-103)."""
    evaluator = R.OneTransEvaluator(trainer.model, trainer.config)
    offline = evaluator.evaluate_offline(data_loader, 'test')
    print('offline:', {k: round(v, 4) for k, v in offline.items() if isinstance(v, float)})
    print('benchmark:', {k: round(v, 3) for k, v in evaluator.benchmark_performance(data_loader, num_batches=5, warmup_batches=2).items()})
    print('report:', evaluator.generate_evaluation_report(data_loader, os.path.join(model_dir, 'evaluation_reports')))


def model_inference_example(trainer, model_dir):
    """Reload the final checkpoint and score one user's candidates, plain and with the per-layer K/V cache
    (OT/examples/train_example.py:106-129; the cached path is north_star item 5)."""
    engine = R.OneTransInferenceEngine(os.path.join(model_dir, 'final_model'))
    cfg = engine.config
    g = torch.Generator().manual_seed(0)
    fc = cfg.feature_config
    mk = lambda names: {n: float(torch.randn((), generator=g)) for n in names}
    user, ctx = mk(fc['user_features']), mk(fc['context_features'])
    seqs = {n: torch.randn(L, cfg.seq_feature_dim, generator=g) for n, L in zip(fc['sequence_features'], (30, 12, 5))}
    items = [mk(fc['item_features']) for _ in range(8)]
    plain = engine.batch_inference([(user, it, ctx, seqs) for it in items])
    cand = {n: [(it[n] if n in it else (user[n] if n in user else ctx[n])) for it in items] for n in cfg.ns_features}
    ranked = engine.rank_candidates(seqs, cand)
    print('ctr, plain batch :', [round(r['ctr'], 4) for r in plain])
    print('ctr, cached user :', [round(v, 4) for v in ranked['ctr']])
    print('stats:', engine.get_stats())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--epochs', type=int, default=3)
    ap.add_argument('--samples', type=int, default=2048)
    ap.add_argument('--model-dir', default='./example_models')
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit('this example needs a CUDA device (sm_100a); the package has no CPU fallback')
    config = example_config()
    basic_training_example(config)
    trainer, data_loader = full_training_pipeline(config, args.epochs, args.samples, args.model_dir)
    evaluation_example(trainer, data_loader, args.model_dir)
    model_inference_example(trainer, args.model_dir)


if __name__ == '__main__':
    main()
