#!/usr/bin/env python
"""BASELINE configs[4]: the multi_experiments.py-style sweep - encoder x decoder variants with 2 / 4 / 8 (ViT) or 2 / 3
(Uformer, the reference requires N == L and L in {2, 3}) frequency bands - as train-step timings on this GPU.

The reference's multi_experiments.py (lines 21-38) only formats one `python train.py ...` command line per experiment and
hands it to os.system; this is the same loop over the option namespace, timing `steps` graph-replayed train steps per
variant (batch 16 per GPU, 128 x 128, synthetic degradations).  Cells the reference leaves undefined are listed as such.
Under torchrun every rank runs the sweep data-parallel (bucketed all-reduce), rank 0 prints.

  python tools/sweep.py [--steps 5] [--batch 16]            -> one JSON object on stdout (rank 0)
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

PKG = bench.PKG

VARIANTS = [
    # name, option overrides, input batch kind
    ('Uformer+Uformer all_3_bands (L=3)', dict(), 'noisy'),
    ('Uformer+Uformer all_2_bands (L=2)', dict(L=2, degradation_embedding_method=['all_2_bands']), 'noisy'),
    ('Uformer+Uformer all_DC (L=3)', dict(degradation_embedding_method=['all_DC']), 'noisy'),
    ('Uformer(origin MSA)+Uformer all_3_bands', dict(encoder_msa_type='origin'), 'noisy'),
    ('ViT 2_bands + DGRN', dict(encoder_type='ViT', decoder_type='ResNet', encoder_dim=64, frequency_decompose_type='2_bands'), 'mixed'),
    ('ViT 4_bands + DGRN', dict(encoder_type='ViT', decoder_type='ResNet', encoder_dim=64, frequency_decompose_type='4_bands'), 'mixed'),
    ('ViT 8_bands + DGRN', dict(encoder_type='ViT', decoder_type='ResNet', encoder_dim=64, frequency_decompose_type='8_bands'), 'mixed'),
    ('ResNet + DGRN', dict(encoder_type='ResNet', decoder_type='ResNet', encoder_dim=256), 'mixed'),
    ('Uformer + DGRN (package-defined adapter)', dict(decoder_type='ResNet'), 'mixed'),
]
UNDEFINED = {
    'Uformer+Uformer all_4_bands / all_8_bands': 'reference: option.py:59-64 asserts L in {2, 3} and decoder_Uformer.py:279-280 indexes all_inter[N]',
    'ViT / ResNet encoder + Uformer decoder': 'reference: the Uformer decoder reads the per-band token tuple only the Uformer encoder produces (decoder_Uformer.py:1124)',
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--only', default=None)
    a = ap.parse_args()
    bench.claim_stdout()
    rank, world, local = (int(os.environ.get(k, '0')) for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'))
    world = max(world, 1)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    synth = importlib.import_module(PKG + '.synth')
    model = importlib.import_module(PKG + '.net.model')
    trainer = importlib.import_module(PKG + '.trainer')
    rows = []
    for name, kw, kind in VARIANTS:
        if a.only and a.only not in name:
            continue
        opt = bench.make_opt(a.batch)
        opt.__dict__.update(kw)
        torch.manual_seed(0)
        try:
            net = model.AirNet(opt).cuda().train()
            ts = trainer.TrainStep(net, lr=2e-4, contrast_loss_weight=0.6 if opt.L == 3 else 0.2, distributed=world > 1)
            x = synth.mixed_batch(a.batch, seed=1234 + 97 * rank) if kind == 'mixed' else synth.noisy_batch(a.batch, 25, seed=1234 + 97 * rank)
            x = [t.cuda() for t in x]
            for _ in range(2):
                ts.step(*x)
            ts.capture(*x)
            ts.step(*x)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                ts.step(*x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            if dist is not None:
                t = torch.tensor([ms], device='cuda')
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = t.item()
            rows.append({'variant': name, 'ms_per_step': ms, 'crops_per_s': world * a.batch / (ms * 1e-3),
                         'loss': float(ts.last['loss']), 'params_M': sum(p.numel() for p in net.parameters()) / 1e6})
            ts.graph = None
            del ts, net
            torch.cuda.empty_cache()
        except Exception as e:                               # a variant the package does not offer is reported, not hidden
            rows.append({'variant': name, 'error': f'{type(e).__name__}: {str(e)[:160]}'})
    if rank == 0:
        bench.emit({'sweep': rows, 'undefined_in_reference': UNDEFINED, 'n_gpus': world, 'batch_per_gpu': a.batch, 'steps': a.steps,
                    'note': 'configs[4]: train-step time of every encoder x decoder x band-count variant the path admits'})
    if dist is not None:
        bench.shutdown(dist)


if __name__ == '__main__':
    main()
