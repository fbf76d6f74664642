#!/usr/bin/env python
"""bench.py - headline benchmark of the restoration-network train step on B200 (see DESIGN.md section "Measurement").

  python bench.py --gpus N --steps K --warmup W            (N>1: launched per rank by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload at N=1 = BASELINE.json configs[1]: Uformer encoder + Uformer decoder (``all_3_bands``, L=3, freq MSA),
full training step of train.py:80-96 (forward q/k encoders + decoder, CE + L1, backward, Adam), 128x128 crops,
batch 16 per GPU, synthetic sigma=25 noisy crops, random-init weights.  One "step" = one such optimisation step.
Prints ONE JSON line (rank 0).  ``value`` = crops/s with inputs resident in HBM; ``e2e`` = the same step driven
from pinned host buffers (H2D of the four crop tensors + D2H of the loss inside the timed region).
``--impl reference`` times the reference algorithm's CPU restatement (oracle/) on the host cores on a bounded
sample of the same workload (the reference itself is pure PyTorch; its modules are not importable on the GPU box).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = 'frequency-wised_all-in-one_image_restoration_model_b200'
METRIC = 'train crops/sec (128x128, Uformer+Uformer all_3_bands train step)'
UNIT = 'crops/s'
BATCH = 16
# dense FLOPs of one train step per crop: 3*34.7 + 34.7 + 3*138.7 GFLOP
STEP_GFLOP_PER_CROP = 555.0


def make_opt(batch, workload='train'):
    o = types.SimpleNamespace(encoder_type='Uformer', decoder_type='Uformer', encoder_dim=256, L=3,
                              encoder_msa_type='freq', encoder_embed_dim=28, embed_dim=56,
                              degradation_embedding_method=['all_3_bands'], frequency_decompose_type='none',
                              learnable_modulator=False, debug_mode=False, batch_size=batch, out_channels=3,
                              batch_wise_decompose=False)
    if workload == 'vit_dgrn_train':        # the authors' own ViT runs: --encoder_dim 64 (plot_LFS_distribution.py:26)
        o.encoder_type, o.decoder_type, o.encoder_dim, o.frequency_decompose_type = 'ViT', 'ResNet', 64, '4_bands'
    elif workload == 'resnet_dgrn_fwd':
        o.encoder_type, o.decoder_type, o.encoder_dim = 'ResNet', 'ResNet', 256
    elif workload == 'uformer_dgrn':          # configs[3]: the pairing the package defines (net/decoder_DGRN.py adapter)
        o.decoder_type = 'ResNet'
    return o


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(self.rows)}


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1 when
    NCCL_DEBUG is set in the environment), so the real stdout is kept aside for the JSON line and fd 1 is pointed at
    stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(obj):
    claim_stdout()
    _JSON_OUT.write(json.dumps(obj) + '\n')
    _JSON_OUT.flush()


def measure_tf32_peak():
    """Dense TF32 tensor-core throughput of this GPU measured the way MEASURED_PEAKS.json measures bf16
    (8192^3 matmul, best of 5).  Only a roofline denominator - cuBLAS is never on the product path."""
    n = 8192
    a = torch.randn(n, n, device='cuda')
    b = torch.randn(n, n, device='cuda')
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = old
    return 2 * n ** 3 / (best * 1e-3) / 1e12


def _load_by_path(name, relpath):
    """A host-side helper module of the package loaded by FILE PATH (no package import, no libfreqair): the reference arm
    must not import the product."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, PKG, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _oracle_state(spec_name, batch, device):
    """Name-keyed deterministic weights (oracle/detfill.py) for the parameter names / shapes of the reference's own
    modules (tests/golden/spec_*.json, written by tools/make_golden*.py from the unmodified reference); the MoCo queue is
    re-shaped to K = 3 * batch (net/model.py:35)."""
    from oracle import detfill
    spec = json.load(open(os.path.join(ROOT, 'tests', 'golden', spec_name)))
    if 'E.E.queue' in spec:
        spec['E.E.queue'][0][2] = 3 * batch
    return {k: v.to(device) for k, v in detfill.make_state(spec).items()}


def oracle_train_step_time(batch, threads, steps=1, device='cpu', warmup=0, workload='train'):
    """Seconds per train step (best of `steps` after `warmup`) of the oracle (the reference algorithm in plain torch ops) at
    ``batch`` crops: forward q/k/decoder, loss, backward, Adam update.  device='cpu' is the contract's reference arm;
    device='cuda' runs the very same torch program on the GPU through stock cuBLAS / cuFFT / ATen kernels (SURVEY section
    8d: "the reference on the same B200 via stock torch CUDA - the real bar the kernels must beat")."""
    from oracle import airnet as oa
    synth = _load_by_path('freqair_synth', 'synth.py')
    torch.set_num_threads(threads)
    cuda = device != 'cpu'
    vit = workload == 'vit_dgrn_train'
    sd = _oracle_state('spec_airnet_vit_dgrn.json' if vit else 'spec_airnet_uformer_uformer_L3.json', batch, device)
    if cuda:
        torch.set_default_device(device)              # the oracle builds its constant tables with bare factories
    pnames = [k[len('E.E.encoder_q.'):] for k in sd if k.startswith('E.E.encoder_q.')
              and not any(s in k for s in ('running_', 'num_batches', 'relative_position_index', 'mask_freq'))]
    train_keys = [k for k in sd if sd[k].is_floating_point() and not k.startswith('E.E.encoder_k.')
                  and not any(s in k for s in ('running_', 'queue', 'mask_freq'))]
    for k in train_keys:
        sd[k].requires_grad_(True)
    opt = torch.optim.Adam([sd[k] for k in train_keys], lr=2e-4)
    torch.set_default_device('cpu')
    xq, xk, clean = (t.to(device) for t in (synth.mixed_batch(batch) if vit else synth.noisy_batch(batch, 25)))
    if cuda:
        torch.set_default_device(device)
    times = []
    try:
        for i in range(warmup + steps):
            if cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            opt.zero_grad()
            if vit:
                restored, logits, _ = oa.airnet_dgrn_forward(sd, xq, xk, True, encoder='ViT', encoder_dim=64,
                                                             decompose_type='4_bands', param_names=pnames)
            else:
                restored, logits, _ = oa.airnet_uformer_forward(sd, xq, xk, True, param_names=pnames)
            labels = torch.zeros(batch, dtype=torch.long, device=device)
            ce = sum(torch.nn.functional.cross_entropy(l, labels) for l in logits) / len(logits)
            loss = (restored - clean).abs().mean() + 0.6 * ce
            loss.backward()
            opt.step()
            if cuda:
                loss.item()
                torch.cuda.synchronize()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    finally:
        torch.set_default_device('cpu')
    return min(times)


def oracle_dgrn_forward_time(batch, threads, steps=1, warmup=0):
    """Seconds per ResNet-encoder + DGRN eval forward (configs[0]) of the oracle at ``batch`` images on the CPU."""
    from oracle import airnet as oa
    synth = _load_by_path('freqair_synth', 'synth.py')
    torch.set_num_threads(threads)
    se, sdg = _oracle_state('spec_resnet_encoder.json', batch, 'cpu'), _oracle_state('spec_dgrn64.json', batch, 'cpu')
    xq, _, _ = synth.noisy_batch(batch, 25)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            _, _, inter = oa.resnet_encoder_forward(se, '', xq)
            oa.dgrn_forward(sdg, '', xq, inter)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return min(times)


def run_reference_cuda(args):
    """Not the contract's reference arm (that one is the CPU run below): the oracle's torch program on cuda:0 with
    stock kernels, fp32 (allow_tf32 off, the parity-equivalent setting) and with allow_tf32 on."""
    out = {'impl': 'reference-torch-cuda', 'metric': METRIC, 'unit': UNIT, 'n_gpus': 1, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': 'configs[1]: Uformer+Uformer all_3_bands train step, 128x128, sigma=25',
                      'batch': args.batch, 'note': 'oracle/ torch program on cuda:0, stock cuBLAS/cuFFT/ATen kernels, eager'}}
    k = max(1, min(args.steps, 5))
    for name, tf32 in (('fp32', False), ('tf32', True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        t = oracle_train_step_time(args.batch, os.cpu_count() or 1, k, 'cuda', 2)
        out[name] = {'value': args.batch / t, 'ms_per_step': t * 1e3}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    out['value'] = out['fp32']['value']
    out['ms_per_step'] = out['fp32']['ms_per_step']
    emit(out)


WORKLOADS = {
    'train': dict(metric=METRIC, unit=UNIT,
                  name='configs[1]: Uformer+Uformer all_3_bands train step, 128x128, sigma=25'),
    'vit_dgrn_train': dict(metric='train crops/sec (128x128, ViT 4_bands + DGRN train step, mixed degradations)', unit=UNIT,
                           name='configs[2]: ViT encoder (4_bands, encoder_dim 64) + DGRN train step, 128x128, '
                                'sigma 15/25/50 + rain + haze'),
    'resnet_dgrn_fwd': dict(metric='eval forward images/sec (128x128, ResNet encoder + DGRN)', unit='img/s',
                            name='configs[0]: ResNet encoder + DGRN eval forward, 128x128, sigma=25, batch 4'),
}


def run_reference(args, rank, world):
    """The contract's reference arm: the reference algorithm's CPU restatement (oracle/) on the host cores, on a bounded
    sample of the arm's workload (batch 2 of the batch-16 train steps; batch 1 of the batch-4 DGRN forward - CPU time is
    linear in batch), --warmup untimed and --steps timed steps (capped so the run ends within minutes)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl = WORKLOADS.get(args.workload, WORKLOADS['train'])
    fwd = args.workload == 'resnet_dgrn_fwd'
    sample_b = 1 if fwd else 2
    w, k = max(0, min(args.warmup, 2)), max(1, min(args.steps, 5))
    if fwd:
        t = oracle_dgrn_forward_time(sample_b, cores, k, w)
        what = f'oracle/ CPU restatement of the ResNet+DGRN forward, batch {sample_b}'
    else:
        t = oracle_train_step_time(sample_b, cores, k, 'cpu', w, args.workload if args.workload in WORKLOADS else 'train')
        what = f'oracle/ CPU restatement of the reference train step, batch {sample_b}'
    v = sample_b / t
    line = {'impl': 'reference', 'metric': wl['metric'], 'value': v, 'unit': wl['unit'], 'n_gpus': world, 'steps': k,
            'warmup': w, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['name'],
                       'sample': f'batch {sample_b} (CPU time is linear in batch); best of {k} step(s) after {w} warm-up'},
            'cpu_baseline': {'value': v, 'unit': wl['unit'], 'cores': cores, 'kind': 'port', 'sample': f'{what}, {k} step(s)'},
            'e2e': {'value': v, 'unit': wl['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


def run_inference(args):
    """Per-image latency of tiled inference (test.py:48-71 geometry): a HxW image -> ceil(H/128)*ceil(W/128) tiles ->
    ONE batched eval forward (query encoder trunk + decoder; the contrastive heads the reference computes and discards
    are skipped) -> overlap-averaged reassembly.  e2e includes the H2D copy of the image and the D2H of the result."""
    size = 512 if '512' in args.workload else 1024
    dgrn = args.workload.endswith('_dgrn')
    pair = 'Uformer encoder + DGRN (documented adapter)' if dgrn else 'Uformer+Uformer all_3_bands'
    synth = importlib.import_module(PKG + '.synth')
    model = importlib.import_module(PKG + '.net.model')
    infer = importlib.import_module(PKG + '.infer')
    ops = importlib.import_module(PKG + '.ops')
    torch.manual_seed(0)
    net = model.AirNet(make_opt(16, 'uformer_dgrn' if dgrn else 'train')).cuda().eval()
    img = synth.gaussian_noise(synth.clean_images(1, size, size, seed=4321), 25, 4322).pin_memory()
    dimg = img.cuda()
    W, K = max(args.warmup, 3), max(args.steps, 1)
    if args.no_graph:
        run = lambda x: infer.restore_tiled(net, x)
    else:
        run = infer.GraphedRestorer(net, size, size)
    for _ in range(W):
        run(dimg)
    torch.cuda.synchronize()
    ops.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        run(dimg)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    launches = ops.launch_count() if args.no_graph else run.graph_launches * K     # a replay re-launches the recorded kernels
    out_host = torch.empty(1, 3, size, size).pin_memory()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(K):
        out_host.copy_(run(img if not args.no_graph else img.cuda(non_blocking=True)), non_blocking=True)
        torch.cuda.synchronize()
    f1.record()
    torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1) / K
    ntiles = (size // 128) ** 2
    emit({'metric': f'{size}x{size} inference ms/img ({pair}, {ntiles} tiles of 128x128)',
                      'value': ms, 'unit': 'ms/img', 'n_gpus': 1, 'steps': K, 'warmup': W, 'ms_per_step': ms,
                      'higher_is_better': False, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                      'config': {'workload': f'tiled eval forward of one {size}x{size} sigma=25 image, random init', 'tiles': ntiles,
                                 'launch': 'eager' if args.no_graph else 'cuda_graph'},
                      'e2e': {'value': ms_e2e, 'unit': 'ms/img', 'h2d_bytes_per_step': img.numel() * 4,
                              'd2h_bytes_per_step': img.numel() * 4},
                      'gpu_launches': launches, 'tiles_per_s': ntiles / (ms * 1e-3)})


def run_dgrn_forward(args):
    """configs[0]: ResNet encoder + DGRN (AirNet eval forward, test.py:59) on the sigma = 25 batch of 4 at 128 x 128.
    value = images / s with the batch resident in HBM (CUDA-graph replay); e2e = pinned host batch -> device -> forward ->
    restored batch back to the host.  cpu_baseline = the oracle's forward on the host cores (batch 1, scaled)."""
    ops = importlib.import_module(PKG + '.ops')
    synth = importlib.import_module(PKG + '.synth')
    model = importlib.import_module(PKG + '.net.model')
    B = 4 if args.batch == BATCH else args.batch
    torch.manual_seed(0)
    net = model.AirNet(make_opt(B, 'resnet_dgrn_fwd')).cuda().eval()
    xq, _, _ = synth.noisy_batch(B, 25)
    host = xq.pin_memory()
    static_in = host.cuda()
    W, K = max(args.warmup, 3), max(args.steps, 1)
    with torch.no_grad():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                net(static_in, static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(graph):
            static_out = net(static_in, static_in)
        per_replay = ops.launch_count() - n0
    for _ in range(W):
        graph.replay()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    clocks = sampler.summary()
    out_host = torch.empty_like(host).pin_memory()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(K):
        static_in.copy_(host, non_blocking=True)
        graph.replay()
        out_host.copy_(static_out, non_blocking=True)
        torch.cuda.synchronize()
    f1.record()
    torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1) / K
    roofline = None
    if not args.no_roofline:
        ops.FLOP_COUNTER[0] = 0
        ops.prof_begin(ops.K_GEMM)
        with torch.no_grad():
            net(static_in, static_in)
        torch.cuda.synchronize()
        gemm_ms, gemm_n = ops.prof_end()
        flops, ops.FLOP_COUNTER[0] = ops.FLOP_COUNTER[0], None
        peak = measure_tf32_peak()
        ach = flops / (gemm_ms * 1e-3) / 1e12
        roofline = {'bound': 'tensor', 'kernel': 'fa_gemm (3x3 / 1x1 / DCN contractions of the forward)', 'achieved': ach,
                    'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak, 'traffic': None, 'launches_per_step': gemm_n,
                    'avg_launch_ms': gemm_ms / max(gemm_n, 1), 'algorithmic_gflop_per_step': flops / 1e9,
                    'share_of_step': gemm_ms / ms, 'peak_source': 'dense TF32 cuBLAS 8192^3 measured in this run'}
    cpu_baseline = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        t = oracle_dgrn_forward_time(1, cores, 1, 0)
        cpu_baseline = {'value': 1 / t, 'unit': 'img/s', 'cores': cores, 'kind': 'port',
                        'sample': f'oracle/ CPU restatement of the same forward at batch 1 ({t:.1f} s; CPU time is linear in batch)'}
    wl = WORKLOADS['resnet_dgrn_fwd']
    emit({'metric': wl['metric'], 'value': B / (ms * 1e-3), 'unit': 'img/s', 'n_gpus': 1, 'steps': K, 'warmup': W,
          'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
          'data': 'synthetic', 'config': {'workload': wl['name'] + ', random init', 'batch': B, 'launch': 'cuda_graph',
                                          'l2': 'one forward streams ~6 GB of activations, far beyond the 126 MB L2'},
          'clocks': clocks,
          'e2e': {'value': B / (ms_e2e * 1e-3), 'unit': 'img/s', 'h2d_bytes_per_step': host.numel() * 4,
                  'd2h_bytes_per_step': host.numel() * 4, 'ms_per_step': ms_e2e},
          'gpu_launches': per_replay * K, 'gpu_launches_per_step': per_replay, 'roofline': roofline,
          'cpu_baseline': cpu_baseline})


def check_grads(net, ts, synth, B, rank, world, dist):
    """with the batch sharded over `world` ranks, the bucketed all-reduce must leave on every rank
    the gradients one GPU computes for the concatenated batch.  Batch-coupled pieces are taken out so that the
    comparison is exact up to summation order: the encoder runs in eval mode (BatchNorm running statistics, no MoCo
    queue), DropPath is off, the loss is the L1 term (mean over the local samples; the mean over ranks is the 1/world
    the optimiser applies).  Both passes go through the same hooks and NCCL buckets: pass 1 with each rank's own shard,
    pass 2 with the full batch on every rank (its all-reduce sums `world` identical copies)."""
    losses = importlib.import_module(PKG + '.losses')
    net.E.eval()
    net.R.train()
    for m in net.modules():
        if hasattr(m, 'drop_path_prob'):
            m.drop_path_prob = 0.0
    b = max(B // world, 1)
    xq, _, clean = synth.noisy_batch(b * world, 25, seed=4242)           # the same global batch on every rank
    xq, clean = xq.cuda(), clean.cuda()

    def grads(x, c):
        ts.zero_grad()
        _, inter = net.E(x, x)
        restored = net.R(x, inter)
        losses.l1_loss(restored, c).backward()
        if ts.ddp is not None:
            ts.ddp.finish()
        torch.cuda.synchronize()
        return [s.grad.clone() / world for s in ts.segments]
    sl = slice(rank * b, (rank + 1) * b)
    g_ddp = grads(xq[sl].contiguous(), clean[sl].contiguous())
    g_one = grads(xq, clean)
    rep = {}
    for name, a, r in zip(('encoder', 'decoder'), g_ddp, g_one):
        scale = r.abs().max().item()
        rep[name] = {'max_abs_err': (a - r).abs().max().item(), 'grad_max_abs': scale, 'n': a.numel(),
                     'rel_l2': ((a - r).norm() / r.norm().clamp_min(1e-30)).item()}
    # worst parameters by name (diagnosis when the check fails)
    worst = []
    for seg, a, r in zip(ts.segments, g_ddp, g_one):
        names = {id(p): n for n, p in net.named_parameters()}
        for p_, o in zip(seg.params, seg.offs):
            e = (a[o:o + p_.numel()] - r[o:o + p_.numel()]).abs().max().item()
            worst.append((e, names.get(id(p_), '?'), o))
    worst.sort(reverse=True)
    rep['worst_parameters'] = [(f'{e:.3e}', n, o) for e, n, o in worst[:12]]
    ok = all(v['max_abs_err'] <= 1e-3 * max(v['grad_max_abs'], 1.0) for k, v in rep.items() if k != 'worst_parameters')
    if dist is not None:
        flag = torch.tensor([1.0 if ok else 0.0], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item() > 0.5)
    if rank == 0:
        emit({'check_grads': rep, 'n_gpus': world, 'global_batch': b * world, 'ok': ok, 'tolerance': '1e-3 max-abs',
              'buckets': len(ts.ddp.buckets) if ts.ddp is not None else 0})
    if not ok:
        raise SystemExit(f'bench.py --check-grads: all-reduced gradients differ from the single-GPU gradients: {rep}')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='freqair', choices=['freqair', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH, help='crops per GPU (the headline config uses 16)')
    ap.add_argument('--ref-device', default='cpu', choices=['cpu', 'cuda'],
                    help='with --impl reference: cpu = the contract arm (default); cuda = the same oracle torch program on '
                         'cuda:0 through stock torch kernels (informational: the stock-library bar on this GPU)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    ap.add_argument('--check-grads', action='store_true',
                    help='multi-GPU parity instead of timing (SURVEY section 8e): the all-reduced gradients of a batch '
                         'sharded over the ranks against the gradients of the same batch on one GPU; prints one JSON line')
    ap.add_argument('--no-graph', action='store_true', help='launch the step kernel by kernel instead of replaying its CUDA graph')
    ap.add_argument('--workload', default='train',
                    choices=['train', 'infer512', 'infer1024', 'infer512_dgrn', 'infer1024_dgrn', 'vit_dgrn_train',
                             'resnet_dgrn_fwd'],
                    help='train = the headline configs[1] step (default); vit_dgrn_train = configs[2] (ViT 4_bands + DGRN '
                         'train step on the mixed-degradation batch, batch-sharded over --gpus); resnet_dgrn_fwd = configs[0] '
                         '(ResNet encoder + DGRN eval forward, batch 4); infer512 / infer1024 = tiled full-resolution inference '
                         'latency, ms per image, 1 GPU, Uformer encoder + Uformer decoder; infer512_dgrn / infer1024_dgrn = '
                         'configs[3]: the same with the DGRN restorer (Uformer encoder + DGRN through the documented adapter)')
    args = ap.parse_args()
    claim_stdout()
    if os.environ.get('FREQAIR_WATCHDOG'):                 # debugging aid: dump every thread's stack and exit if stuck
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['FREQAIR_WATCHDOG']), exit=True)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        if args.ref_device == 'cuda':
            if rank == 0:
                run_reference_cuda(args)
            return
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device - the freqair path has no CPU fallback (use --impl reference for the CPU baseline)')
    if args.workload.startswith('infer'):
        run_inference(args)
        return
    if args.workload == 'resnet_dgrn_fwd':
        run_dgrn_forward(args)
        return
    vit = args.workload == 'vit_dgrn_train'
    wl = WORKLOADS[args.workload]
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # The gradient all-reduce needs ~10 GB/s (1.09 GB per 110 ms step) and overlaps persistent one-CTA-per-SM
        # contraction kernels: every SM NCCL holds costs those kernels a wave.  Measured at 2 GPUs (ms/step): NCCL default
        # 113.2, 16 CTAs 111.6, 8 CTAs 113.4, 4 CTAs 114.8 (too few: the last bucket's tail shows) - 16 unless the user says otherwise.
        os.environ.setdefault('NCCL_MAX_CTAS', '16')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ops = importlib.import_module(PKG + '.ops')
    synth = importlib.import_module(PKG + '.synth')
    model = importlib.import_module(PKG + '.net.model')
    trainer = importlib.import_module(PKG + '.trainer')
    B = args.batch
    torch.manual_seed(0)                                   # identical replicas on every rank
    net = model.AirNet(make_opt(B, args.workload)).cuda().train()
    ts = trainer.TrainStep(net, lr=3e-4 if vit else 2e-4, contrast_loss_weight=0.6, distributed=world > 1)   # option.py:80-103
    if vit:                                                                # sample i cycles sigma15/25/50, rain, haze
        xq, xk, clean = synth.mixed_batch(B, seed=1234 + 97 * rank)
    else:
        xq, xk, clean = synth.noisy_batch(B, 25, seed=1234 + 97 * rank)   # different crops per rank
    host = [t.pin_memory() for t in (xq, xk, clean, clean.clone())]       # train.py:77-78 copies four tensors
    dev_in = [t.cuda(non_blocking=True) for t in host[:3]]
    h2d_bytes = sum(t.numel() * 4 for t in host)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.check_grads:
        check_grads(net, ts, synth, B, rank, world, dist)
        if dist is not None:
            shutdown(dist, ts)
        return
    for _ in range(W):
        ts.step(*dev_in)
    barrier()
    if not args.no_graph:
        ts.capture(*dev_in)                                # whole step -> one CUDA graph; step() replays it from here on
        for _ in range(2):
            ts.step(*dev_in)
        barrier()
    # ---------------------------------------------------------------- value: device-resident inputs
    sampler = ClockSampler(local)
    sampler.start()
    ops.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        ts.step(*dev_in)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count()
    if ts.graph is not None:                               # a replay re-launches every kernel recorded at capture
        launches = ts.graph_launches * K
    clocks = sampler.summary()
    # ---------------------------------------------------------------- e2e: pinned host -> device every step, loss read back
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = 0.0
    for _ in range(K):
        d = [t.cuda(non_blocking=True) for t in host]
        loss = ts.step(d[0], d[1], d[2])
        last = float(loss.item())                          # D2H of the step's result (4 bytes) - also a sync, as train.py:98 logging does
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    if dist is not None:
        tt = torch.tensor([ms, ms_e2e], device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = tt.tolist()
    value = world * B * K / (ms * 1e-3)
    e2e = world * B * K / (ms_e2e * 1e-3)

    roofline = roofline_hbm = None
    cpu_baseline = None
    if not args.no_roofline:
        # dominant kernel class = the dense contractions (fa_gemm): time every launch with in-stream events
        # (kernel-by-kernel launch of the same step: a graph replay cannot be bracketed per launch).
        # Every rank runs this step (it contains the gradient all-reduce); rank 0 reports.
        ts.graph = None
        ops.FLOP_COUNTER[0] = 0
        ops.prof_begin(ops.K_GEMM)
        ts.step(*dev_in)
        torch.cuda.synchronize()
        gemm_ms, gemm_n = ops.prof_end()
        flops = ops.FLOP_COUNTER[0]
        ops.FLOP_COUNTER[0] = None
        peak = measure_tf32_peak() if rank == 0 else 1.0
        ach = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        # DRAM bytes per launch of the same kernel class from the committed ncu pass over one step (tools/one_step.py
        # under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`; profiles/r2_step_traffic.json)
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, 'profiles', 'r2_step_traffic.json')
        if os.path.exists(tpath) and not vit:
            tj = json.load(open(tpath))
            traffic = tj.get('gemm', {}).get('dram_bytes_per_launch')
            traffic_note = tj.get('note')
        roofline = {'bound': 'tensor', 'kernel': 'fa_gemm (all dense contractions of the step)', 'achieved': ach,
                    'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak if peak else None, 'traffic': traffic,
                    'traffic_note': traffic_note,
                    'launches_per_step': gemm_n, 'avg_launch_ms': gemm_ms / max(gemm_n, 1),
                    'algorithmic_gflop_per_step': flops / 1e9, 'share_of_step': gemm_ms / (ms / K),
                    'share_basis': 'sum of the per-launch event times (kernel-by-kernel launch, every kernel alone on the '
                                   'GPU) over the graph-replay step time; the ncu launch list in profiles/ gives the share '
                                   'under serialised launches',
                    'peak_source': 'dense TF32 cuBLAS 8192^3 measured in this run (MEASURED_PEAKS.json holds bf16 only: '
                                   'the contractions run fp32/tf32, SURVEY.md section 8d)',
                    'note': 'achieved counts ALGORITHMIC flops (2MNK) of ~1 000 launches of ~100 shapes; the restorer\'s LeFF '
                            'contractions run 1xTF32, every other layer class 3xTF32 (3 MMAs per product, needed for the '
                            '1e-3 parity bar: DESIGN.md section 3), and about half of the launches are HBM-class shapes '
                            '(K <= 224 at the 128^2 / 64^2 levels; per-shape GB/s in profiles/r2_bench_kernels_gemm.txt)'}
        # second roofline: the HBM-bound class (depthwise conv + LayerNorm), algorithmic bytes / event time
        hbm_peak = None
        ppath = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(ppath):
            hbm_peak = json.load(open(ppath)).get('hbm_gbs')
        bw_ms = bw_n = 0
        ops.BYTE_COUNTER[0] = 0
        for cls in (ops.K_DWCONV, ops.K_LN):
            ops.prof_begin(cls)
            ts.step(*dev_in)
            torch.cuda.synchronize()
            m_, n_ = ops.prof_end()
            bw_ms += m_
            bw_n += n_
        nbytes = ops.BYTE_COUNTER[0] / 2            # both classes were counted in both profiled steps
        ops.BYTE_COUNTER[0] = None
        roofline_hbm = {'bound': 'hbm', 'kernel': 'fa_dwconv3x3_fwd/bwd + fa_layernorm_fwd/bwd', 'achieved': nbytes / (bw_ms * 1e-3) / 1e9 if bw_ms else 0.0,
                        'peak': hbm_peak if hbm_peak else 6650.0, 'unit': 'GB/s',
                        'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if hbm_peak else '6.65 TB/s (of fallback)',
                        'launches_per_step': bw_n, 'avg_launch_ms': bw_ms / max(bw_n, 1),
                        'algorithmic_gb_per_step': nbytes / 1e9, 'share_of_step': bw_ms / (ms / K), 'traffic': None}
        roofline_hbm['frac'] = roofline_hbm['achieved'] / roofline_hbm['peak']
        if os.path.exists(tpath) and not vit:
            roofline_hbm['traffic'] = json.load(open(tpath)).get('dwconv_ln', {}).get('dram_bytes_per_launch')
    if dist is not None:
        dist.barrier()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sb = 2
        t = oracle_train_step_time(sb, cores, 3, 'cpu', 1, args.workload)   # 1 warm-up (allocator, thread pool), best of 3
        cpu_baseline = {'value': sb / t, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                        'sample': f'oracle/ CPU restatement of the same train step at batch {sb} (best of 3 steps after 1 '
                                  f'warm-up, {t:.1f} s/step; CPU step time is linear in batch)'}
    if rank == 0:
        line = {'metric': wl['metric'], 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
                'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': wl['name'] + '; full train step incl. Adam, random init',
                           'batch_per_gpu': B, 'global_batch': B * world, 'parallelism': f'dp{world}',
                           'launch': 'eager' if args.no_graph else 'cuda_graph',
                           'l2': 'no explicit flush: one step streams >20 GB of activations + 4.5 GB of weights/optimizer '
                                 'state, far beyond the 126 MB L2'},
                'clocks': clocks,
                'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
                        'ms_per_step': ms_e2e / K, 'last_loss': last},
                'gpu_launches': launches, 'gpu_launches_per_step': launches / K,
                'step_tflop': (STEP_GFLOP_PER_CROP * B / 1e3) if not vit else None,
                'roofline': roofline, 'roofline_hbm': roofline_hbm, 'cpu_baseline': cpu_baseline}
        emit(line)
    if dist is not None:
        shutdown(dist, ts)


def shutdown(dist, ts=None):
    """Leave a multi-rank run for certain.  A captured CUDA graph that contains NCCL kernels must be released before the
    communicator is destroyed (with `--no-roofline` the graph used to stay alive and destroy_process_group never returned:
    an 8-GPU run printed its line and then sat until the harness killed it).  The JSON line is out by now, so after a last
    barrier every rank drops the graph, destroys the group under a watchdog and exits the interpreter without running
    further destructors."""
    if ts is not None:
        ts.graph = None
        ts.static_out = ts.static_in = None
    torch.cuda.synchronize()
    try:
        dist.barrier()
    except Exception:
        pass
    wd = threading.Timer(60.0, lambda: os._exit(0))            # if teardown hangs anyway: leave after 60 s (the line is out)
    wd.daemon = True
    wd.start()
    try:
        dist.destroy_process_group()
    except Exception:
        pass
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == '__main__':
    main()
