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


def make_opt(batch):
    return types.SimpleNamespace(encoder_type='Uformer', decoder_type='Uformer', encoder_dim=256, L=3,
                                 encoder_msa_type='freq', encoder_embed_dim=28, embed_dim=56,
                                 degradation_embedding_method=['all_3_bands'], frequency_decompose_type='none',
                                 learnable_modulator=False, debug_mode=False, batch_size=batch, out_channels=3,
                                 batch_wise_decompose=False)


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


def oracle_train_step_time(batch, threads, steps=1, device='cpu'):
    """Seconds per train step of the oracle (reference algorithm in plain torch ops) at ``batch`` crops: forward
    q/k/decoder, loss, backward, Adam update.  device='cpu' is the contract's reference arm; device='cuda' runs the very
    same torch program on the GPU through stock cuBLAS / cuFFT / ATen kernels ("the reference on
    the same B200 via stock torch CUDA - the real bar the kernels must beat")."""
    from oracle import airnet as oa
    synth = importlib.import_module(PKG + '.synth')
    model = importlib.import_module(PKG + '.net.model')
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = model.AirNet(make_opt(batch))               # parameter container only (CPU); the math below is oracle/
    sd = {k: v.detach().clone().to(device) for k, v in net.state_dict().items()}
    del net
    cuda = device != 'cpu'
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
    xq, xk, clean = (t.to(device) for t in synth.noisy_batch(batch, 25))
    if cuda:
        torch.set_default_device(device)
    times = []
    try:
        for _ in range(steps):
            if cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            opt.zero_grad()
            restored, logits, _ = oa.airnet_uformer_forward(sd, xq, xk, True, param_names=pnames)
            labels = torch.zeros(batch, dtype=torch.long, device=device)
            ce = sum(torch.nn.functional.cross_entropy(l, labels) for l in logits) / len(logits)
            loss = (restored - clean).abs().mean() + 0.6 * ce
            loss.backward()
            opt.step()
            if cuda:
                loss.item()
                torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    finally:
        torch.set_default_device('cpu')
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
        oracle_train_step_time(args.batch, os.cpu_count() or 1, 2, 'cuda')
        t = oracle_train_step_time(args.batch, os.cpu_count() or 1, k, 'cuda')
        out[name] = {'value': args.batch / t, 'ms_per_step': t * 1e3}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    out['value'] = out['fp32']['value']
    out['ms_per_step'] = out['fp32']['ms_per_step']
    emit(out)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_b = 2
    for _ in range(max(0, min(args.warmup, 1))):
        oracle_train_step_time(sample_b, cores, 1)
    k = max(1, min(args.steps, 3))
    t = oracle_train_step_time(sample_b, cores, k)
    v = sample_b / t
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': world, 'steps': k,
            'warmup': min(args.warmup, 1), 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'configs[1]: Uformer+Uformer all_3_bands train step, 128x128, sigma=25',
                       'sample': f'batch {sample_b} of the batch-{BATCH} step (CPU step time is linear in batch)'},
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': f'oracle/ CPU restatement of the reference train step, batch {sample_b}, {k} step(s)'},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


def run_inference(args):
    """Per-image latency of tiled inference (test.py:48-71 geometry): a HxW image -> ceil(H/128)*ceil(W/128) tiles ->
    ONE batched eval forward (query encoder trunk + decoder; the contrastive heads the reference computes and discards
    are skipped) -> overlap-averaged reassembly.  e2e includes the H2D copy of the image and the D2H of the result."""
    size = 512 if args.workload == 'infer512' else 1024
    synth = importlib.import_module(PKG + '.synth')
    model = importlib.import_module(PKG + '.net.model')
    infer = importlib.import_module(PKG + '.infer')
    ops = importlib.import_module(PKG + '.ops')
    torch.manual_seed(0)
    net = model.AirNet(make_opt(16)).cuda().eval()
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
    launches = ops.launch_count()
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
    emit({'metric': f'{size}x{size} inference ms/img (Uformer+Uformer all_3_bands, {ntiles} tiles of 128x128)',
                      'value': ms, 'unit': 'ms/img', 'n_gpus': 1, 'steps': K, 'warmup': W, 'ms_per_step': ms,
                      'higher_is_better': False, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                      'config': {'workload': f'tiled eval forward of one {size}x{size} sigma=25 image, random init', 'tiles': ntiles,
                                 'launch': 'eager' if args.no_graph else 'cuda_graph'},
                      'e2e': {'value': ms_e2e, 'unit': 'ms/img', 'h2d_bytes_per_step': img.numel() * 4,
                              'd2h_bytes_per_step': img.numel() * 4},
                      'gpu_launches': launches, 'tiles_per_s': ntiles / (ms * 1e-3)})


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
    ap.add_argument('--no-graph', action='store_true', help='launch the step kernel by kernel instead of replaying its CUDA graph')
    ap.add_argument('--workload', default='train', choices=['train', 'infer512', 'infer1024'],
                    help='train = the headline configs[1] step (default); infer512 / infer1024 = configs[3]-style tiled '
                         'full-resolution inference latency (Uformer encoder + Uformer decoder), ms per image, 1 GPU')
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
    if args.workload != 'train':
        run_inference(args)
        return
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ops = importlib.import_module(PKG + '.ops')
    synth = importlib.import_module(PKG + '.synth')
    model = importlib.import_module(PKG + '.net.model')
    trainer = importlib.import_module(PKG + '.trainer')
    B = args.batch
    torch.manual_seed(0)                                   # identical replicas on every rank
    net = model.AirNet(make_opt(B)).cuda().train()
    ts = trainer.TrainStep(net, lr=2e-4, contrast_loss_weight=0.6, distributed=world > 1)
    xq, xk, clean = synth.noisy_batch(B, 25, seed=1234 + 97 * rank)       # different crops per rank
    host = [t.pin_memory() for t in (xq, xk, clean, clean.clone())]       # train.py:77-78 copies four tensors
    dev_in = [t.cuda(non_blocking=True) for t in host[:3]]
    h2d_bytes = sum(t.numel() * 4 for t in host)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

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

    roofline = None
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
        roofline = {'bound': 'tensor', 'kernel': 'fa_gemm (all dense contractions of the step)', 'achieved': ach,
                    'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak if peak else None, 'traffic': None,
                    'launches_per_step': gemm_n, 'avg_launch_ms': gemm_ms / max(gemm_n, 1),
                    'algorithmic_gflop_per_step': flops / 1e9, 'share_of_step': gemm_ms / (ms / K),
                    'share_basis': 'sum of the per-launch event times (kernel-by-kernel launch, every kernel alone on the '
                                   'GPU) over the graph-replay step time; the ncu launch list in profiles/ gives 58 %',
                    'peak_source': 'dense TF32 cuBLAS 8192^3 measured in this run (MEASURED_PEAKS.json holds bf16 only: '
                                   'the contractions run fp32/tf32, SURVEY.md section 8d)',
                    'note': 'achieved counts ALGORITHMIC flops (2MNK); the product path issues 3 tf32 MMAs per product '
                            '(error-compensated 3xTF32, needed for the 1e-3 parity bar), so the tensor pipe does 3x this '
                            'work (compute-class shapes run 175-198 TFLOP/s algorithmic = 520-590 TFLOP/s of raw tf32 MMA, DESIGN.md section 3); '
                            'K <= 224 layers at the 128^2 / 64^2 levels are HBM-bound (2.3-3.2 TB/s, tools/bench_kernels.py)'}
    if dist is not None:
        dist.barrier()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sb = 2
        oracle_train_step_time(sb, cores, 1)                 # warm-up (allocator, thread pool)
        t = oracle_train_step_time(sb, cores, 3)
        cpu_baseline = {'value': sb / t, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                        'sample': f'oracle/ CPU restatement of the same train step at batch {sb} (best of 3 steps after 1 '
                                  f'warm-up, {t:.1f} s/step; CPU step time is linear in batch)'}
    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
                'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': 'configs[1]: Uformer encoder + Uformer decoder (all_3_bands, L=3, freq MSA) full '
                                       'train step incl. Adam, 128x128 crops, sigma=25 synthetic noise, random init',
                           'batch_per_gpu': B, 'global_batch': B * world, 'parallelism': f'dp{world}',
                           'launch': 'eager' if args.no_graph else 'cuda_graph',
                           'l2': 'no explicit flush: one step streams >20 GB of activations + 4.5 GB of weights/optimizer '
                                 'state, far beyond the 126 MB L2'},
                'clocks': clocks,
                'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
                        'ms_per_step': ms_e2e / K, 'last_loss': last},
                'gpu_launches': launches, 'gpu_launches_per_step': launches / K,
                'step_tflop': STEP_GFLOP_PER_CROP * B / 1e3,
                'roofline': roofline, 'cpu_baseline': cpu_baseline}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
