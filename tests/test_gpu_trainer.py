"""GPU: the fused train step (trainer.TrainStep) - direct gradient sinks vs gradients routed through autograd,
CUDA-graph replay vs kernel-by-kernel launch, and Adam vs torch.optim.Adam on the same gradients."""
import importlib
import types

import pytest
import torch

from conftest import PKG_NAME

pytestmark = pytest.mark.gpu


def make_opt(batch):
    return types.SimpleNamespace(encoder_type='Uformer', decoder_type='Uformer', encoder_dim=256, L=3,
                                 encoder_msa_type='freq', encoder_embed_dim=28, embed_dim=56,
                                 degradation_embedding_method=['all_3_bands'], frequency_decompose_type='none',
                                 learnable_modulator=False, debug_mode=False, batch_size=batch, out_channels=3,
                                 batch_wise_decompose=False)


def build(batch=2):
    model = importlib.import_module(PKG_NAME + '.net.model')
    trainer = importlib.import_module(PKG_NAME + '.trainer')
    synth = importlib.import_module(PKG_NAME + '.synth')
    torch.manual_seed(0)
    net = model.AirNet(make_opt(batch)).cuda().train()
    for m in net.modules():                      # DropPath off: the two runs must see identical graphs
        if hasattr(m, 'drop_path_prob'):
            m.drop_path_prob = 0.0
    ts = trainer.TrainStep(net, lr=2e-4, contrast_loss_weight=0.6)
    x = [t.cuda() for t in synth.noisy_batch(batch, 25)]
    return net, ts, x


def test_direct_grad_equals_autograd_grad():
    lewin = importlib.import_module(PKG_NAME + '.net.lewin')
    net, ts, x = build()
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    grads = {}
    for direct in (True, False):
        net.load_state_dict(sd0)
        lewin.DIRECT_GRAD = direct
        ts.zero_grad()
        restored, logits, labels = net(*x[:2])
        loss, _, _ = ts.loss(restored, logits, labels, x[2])
        loss.backward()
        grads[direct] = [s.grad.clone() for s in ts.segments]
    lewin.DIRECT_GRAD = True
    for a, b in zip(grads[True], grads[False]):
        scale = b.abs().max().item()
        assert (a - b).abs().max().item() <= 2e-5 * scale + 1e-8
        assert a.abs().sum().item() > 0


def test_graph_replay_matches_eager_and_adam_matches_torch():
    net, ts, x = build()
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    # eager: three steps
    eager_losses = [ts.step(*x).item() for _ in range(3)]
    flat_eager = [s.flat.clone() for s in ts.segments]
    # torch.optim.Adam on the same gradients (first step only: afterwards the parameters differ by round-off)
    net.load_state_dict(sd0)
    net2_params = [p for s in ts.segments for p in s.params]
    ref = [p.detach().clone() for p in net2_params]
    ts.t = 0
    for s in ts.segments:
        s.m.zero_(); s.v.zero_()
    ts.step(*x)
    gr = [p.grad.detach().clone() for p in net2_params]
    ref_p = [r.clone().requires_grad_(True) for r in ref]
    opt = torch.optim.Adam(ref_p, lr=2e-4)
    for r, g in zip(ref_p, gr):
        r.grad = g
    opt.step()
    worst = max((p.detach() - r.detach()).abs().max().item() for p, r in zip(net2_params, ref_p))
    assert worst < 1e-6, worst
    # graph: restore the initial state, capture, replay three steps
    net.load_state_dict(sd0)
    ts.t = 0
    for s in ts.segments:
        s.m.zero_(); s.v.zero_()
    net.E.E.queue_ptr.zero_()
    ts.capture(*x, warmup=0)
    graph_losses = [ts.step(*x).item() for _ in range(3)]
    assert ts.graph_launches > 1000
    for a, b in zip(eager_losses, graph_losses):
        assert abs(a - b) <= 2e-4 * max(abs(a), 1.0), (eager_losses, graph_losses)


def test_encoder_only_phase_and_spectral_l1_term():
    """train.py:84-87 (encoder phase: contrastive loss only; decoder weights and Adam state untouched) and
    train.py:90-91 (spectral L1 term) against torch autograd on the oracle's decompose."""
    from oracle import freq
    trainer = importlib.import_module(PKG_NAME + '.trainer')
    net, ts, x = build()
    ts.encoder_only = True
    dec0 = ts.segments[1].flat.clone()
    enc0 = ts.segments[0].flat.clone()
    l = ts.step(*x)
    assert torch.isfinite(l).all() and ts.last['l1'].item() == 0.0
    assert torch.equal(ts.segments[1].flat, dec0) and ts.segments[1].m.abs().max().item() == 0
    assert not torch.equal(ts.segments[0].flat, enc0)
    assert ts.ts == [1, 0]
    ts.capture(*x, warmup=0)                       # the encoder-only step is graph-capturable as well
    l2 = ts.step(*x)
    assert torch.isfinite(l2).all() and ts.ts == [2, 0] and torch.equal(ts.segments[1].flat, dec0)

    net, _, x = build()
    ts = trainer.TrainStep(net, lr=2e-4, contrast_loss_weight=0.6, num_frequency_bands_l1=4, frequency_l1_loss_weight=0.1)
    restored = torch.rand(2, 3, 128, 128, device='cuda').requires_grad_(True)
    logits = [torch.randn(2, 7, device='cuda') for _ in range(3)]
    labels = [torch.zeros(2, dtype=torch.long, device='cuda') for _ in range(3)]
    loss, l1, ce = ts.loss(restored, logits, labels, x[2])
    loss.backward()
    r = restored.detach().cpu().requires_grad_(True)
    c = x[2].cpu()
    D = lambda t: freq.decompose(t, 'frequency_decompose', 0.25, inverse=False)
    ref = (r - c).abs().mean() + 0.1 * (D(r) - D(c)).abs().mean()
    ref.backward()
    assert abs(l1.item() - ref.item()) <= 1e-5 * abs(ref.item())
    d = (restored.grad.cpu() - r.grad)
    assert d.norm().item() <= 2e-3 * r.grad.norm().item()


def test_folded_droppath_equals_materialised_scale():
    """DropPath backward folded into the contractions (a_kscale / epilogue row scale) against the fa_scale_rows path:
    same DropPath draws (forced), same gradients."""
    lewin = importlib.import_module(PKG_NAME + '.net.lewin')
    net, ts, x = build()
    g = torch.Generator(device='cuda').manual_seed(3)
    for m in net.modules():
        if hasattr(m, 'forced_dp'):
            B = 2 * (3 if 'encoder' in type(m).__module__ else 1)
            draw = lambda: (torch.rand(B, device='cuda', generator=g) > 0.3).float() / 0.7
            m.forced_dp = (draw(), draw())
    grads = {}
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    try:
        for fold in (True, False):
            net.load_state_dict(sd0)                 # the forward moves the MoCo queue, key encoder and BN statistics
            lewin.FOLD_DROPPATH = fold
            ts.zero_grad()
            restored, logits, labels = net(*x[:2])
            loss, _, _ = ts.loss(restored, logits, labels, x[2])
            loss.backward()
            grads[fold] = [s.grad.clone() for s in ts.segments]
    finally:
        lewin.FOLD_DROPPATH = True
    for a, b in zip(grads[True], grads[False]):
        scale = b.abs().max().item()
        assert (a - b).abs().max().item() <= 2e-5 * scale + 1e-8
        assert a.abs().sum().item() > 0


def test_distributed_bookkeeping_single_rank():
    """The data-parallel step on a 1-rank NCCL group: every parameter must be reported complete exactly once per backward
    (block-level nodes report their parameters directly, and the engine STILL runs AccumulateGrad - and its hook - for
    them with an undefined gradient; round 1 counted both, so buckets were reduced before their last gradients landed).
    BucketedAllReduce.finish() raises when a gradient is reported after its bucket was launched; with one rank the
    reduced gradients must also equal the plain ones.  The 2-GPU equality is `bench.py --gpus 2 --check-grads`."""
    import os
    import socket
    import torch.distributed as dist
    trainer = importlib.import_module(PKG_NAME + '.trainer')
    model = importlib.import_module(PKG_NAME + '.net.model')
    synth = importlib.import_module(PKG_NAME + '.synth')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', 0))
    try:
        torch.manual_seed(0)
        net = model.AirNet(make_opt(2)).cuda().train()
        for m in net.modules():
            if hasattr(m, 'drop_path_prob'):
                m.drop_path_prob = 0.0
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        x = [t.cuda() for t in synth.noisy_batch(2, 25)]
        ts = trainer.TrainStep(net, lr=2e-4, contrast_loss_weight=0.6, distributed=True, bucket_mb=8)
        assert len(ts.ddp.buckets) > 20
        ts.zero_grad()
        restored, logits, labels = net(*x[:2])
        loss, _, _ = ts.loss(restored, logits, labels, x[2])
        loss.backward()
        n_early = sum(ts.ddp.launched)
        ts.ddp.finish()                                   # raises on late / double reports
        assert n_early > len(ts.ddp.buckets) // 2         # most buckets were reduced while backward was still running
        g_ddp = [s.grad.clone() for s in ts.segments]
        ts.ddp, ddp = None, ts.ddp                        # the same step without the all-reduce
        for seg in ts.segments:
            for p in seg.params:
                p._fa_ready = None
        net.load_state_dict(sd0)
        ts.zero_grad()
        restored, logits, labels = net(*x[:2])
        loss, _, _ = ts.loss(restored, logits, labels, x[2])
        loss.backward()
        for a, s in zip(g_ddp, ts.segments):
            assert (a - s.grad).abs().max().item() <= 2e-5 * s.grad.abs().max().item() + 1e-8
    finally:
        dist.destroy_process_group()
