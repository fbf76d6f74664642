"""The training step of the reference's train.py (lines 80-96: zero_grad, forward, CE + L1 loss, backward,
Adam) as one object, plus what a single 8xB200 box adds: batch-sharded data parallelism with a bucketed
gradient all-reduce over NCCL/NVLink that overlaps the remaining backward.

Memory layout: every trainable parameter is a view of one flat fp32 buffer per segment (query encoder,
restorer); gradients, Adam m and v mirror that layout, so zero_grad is one memset, Adam is one fused kernel
per segment (train.py:63,96 -> fa_adam_step) and all-reduce buckets are plain slices of the gradient buffer.
"""
import torch
import torch.nn.functional as F

from . import ops
from .losses import l1_loss, spectral_l1_loss
from .net import lewin


def _offsets(params):
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += (p.numel() + 3) // 4 * 4
    return offs, off


def flatten_into(params):
    """Make ``params`` views of one flat buffer (values preserved). Returns (flat, offsets)."""
    params = list(params)
    offs, total = _offsets(params)
    flat = torch.zeros(total, device=params[0].device, dtype=torch.float32)
    for p, o in zip(params, offs):
        v = flat[o:o + p.numel()].view(p.shape)
        v.copy_(p.data)
        p.data = v
    return flat, offs


class Segment:
    """A group of parameters sharing one flat buffer, with matching flat gradient / Adam state."""

    def __init__(self, params, flat=None):
        self.params = list(params)
        offs, total = _offsets(self.params)
        if flat is None or flat.numel() != total or self.params[0].data_ptr() != flat.data_ptr():
            flat, offs = flatten_into(self.params)
        self.flat, self.offs = flat, offs
        self.grad = torch.zeros_like(flat)
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        for p, o in zip(self.params, offs):
            p.grad = self.grad[o:o + p.numel()].view(p.shape)


class BucketedAllReduce:
    """Gradient SUM over ranks in ~bucket_mb slices of the flat gradient buffers, launched from
    post-accumulate-grad hooks on a side stream as soon as every parameter of a bucket has its gradient.
    The 1/world_size of the mean rides the optimiser kernel (``fa_adam_step(grad_scale=...)``), so no extra pass
    touches the 1.09 GB of gradients.  Device-agnostic on purpose: the bucket / hook logic is exercised on CPU
    with the gloo backend at world_size 2 (tests/test_ddp_gloo.py); on CUDA the collectives run on a side stream."""

    def __init__(self, segments, bucket_mb=64, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.on_cuda = segments[0].grad.is_cuda
        self.stream = torch.cuda.Stream() if self.on_cuda else None
        self.world = dist.get_world_size(group)
        self.buckets = []          # (tensor slice, n_params)
        self.bucket_of = {}        # id(param) -> bucket index
        self.seen = set()
        self.pending = []
        self.handles = []
        per = int(bucket_mb * (1 << 20) // 4)
        for seg in segments:
            n = seg.grad.numel()
            nb = max(1, (n + per - 1) // per)
            bounds = [min(n, i * per) for i in range(nb + 1)]
            base = len(self.buckets)
            counts = [0] * nb
            owners = []
            for p, o in zip(seg.params, seg.offs):
                b = min(nb - 1, (o + p.numel() - 1) // per)      # bucket of the parameter's last element
                counts[b] += 1
                owners.append(base + b)
            for i in range(nb):
                self.buckets.append((seg.grad[bounds[i]:bounds[i + 1]], counts[i]))
            for p, b in zip(seg.params, owners):
                p.register_post_accumulate_grad_hook(self._make_hook(b))
                self.bucket_of[id(p)] = b
        self.reset()

    def param_ready(self, p):
        """A block-level backward accumulated this parameter's gradient straight into the flat buffer (no
        AccumulateGrad, so no hook): same bookkeeping as the hook, once per parameter per step."""
        k = id(p)
        b = self.bucket_of.get(k)
        if b is None or k in self.seen:
            return
        self.seen.add(k)
        self.pending[b] -= 1
        if self.pending[b] == 0:
            self._launch(b)

    def reset(self):
        self.seen = set()
        self.pending = [c for _, c in self.buckets]
        self.handles = []
        self.launched = [False] * len(self.buckets)

    def _make_hook(self, b):
        def hook(_param):
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        if self.launched[b]:
            return
        self.launched[b] = True
        if not self.on_cuda:
            self.handles.append(self.dist.all_reduce(self.buckets[b][0], op=self.dist.ReduceOp.SUM, group=self.group,
                                                     async_op=True))
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            self.handles.append(self.dist.all_reduce(self.buckets[b][0], op=self.dist.ReduceOp.SUM, group=self.group,
                                                     async_op=True))

    def finish(self):
        for b, (_, c) in enumerate(self.buckets):
            if not self.launched[b]:           # parameters that received no gradient this step
                self._launch(b)
        for h in self.handles:
            h.wait()
        if self.on_cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.reset()


class TrainStep:
    """The body of the reference's training loop (train.py:80-96).

    ``encoder_only=True`` is the first phase (``epoch < opt.epochs_encoder``, train.py:84-87): contrastive loss on
    ``net.E`` alone; the decoder receives no gradient and - as torch.optim.Adam skips parameters without a gradient -
    neither its weights nor its Adam state / step count move.  ``num_frequency_bands_l1 != -1`` adds the spectral L1
    term of train.py:69-70,90-91 with weight ``frequency_l1_loss_weight``."""

    def __init__(self, net, lr=2e-4, contrast_loss_weight=0.6, betas=(0.9, 0.999), eps=1e-8, distributed=False,
                 bucket_mb=64, encoder_only=False, num_frequency_bands_l1=-1, frequency_l1_loss_weight=0.1,
                 patch_size=128):
        self.net = net
        self.lr, self.betas, self.eps = lr, betas, eps
        self.w = contrast_loss_weight
        self.encoder_only = encoder_only
        self.w_freq = frequency_l1_loss_weight
        self.decompose = None
        if num_frequency_bands_l1 != -1:
            from .net.utils.frequency_decompose import FrequencyDecompose
            self.decompose = FrequencyDecompose('frequency_decompose', 1. / num_frequency_bands_l1, patch_size, patch_size,
                                                inverse=False)                       # train.py:70
        self.ts = [0, 0]                # Adam step count of the encoder / decoder segment
        moco = net.E.E
        fq, _ = moco._ensure_flat()
        enc_params = [p for p in moco.encoder_q.parameters()]
        dec_params = [p for p in net.R.parameters()]
        self.segments = [Segment(enc_params, fq), Segment(dec_params)]
        moco._flat_q = self.segments[0].flat
        self.ddp = BucketedAllReduce(self.segments, bucket_mb) if distributed else None
        # parameters now own .grad views of per-step-zeroed flat buffers: let the block backwards accumulate into them
        lewin.DIRECT_GRAD = True
        lewin.GRAD_READY = self.ddp.param_ready if self.ddp is not None else None
        self.last = {}
        # CUDA-graph state (capture()): static inputs / loss, and the two step-dependent Adam scalars in device memory
        self.graph = None
        self.graph_launches = 0
        self.static_in = None
        self.static_out = None
        dev = self.segments[0].flat.device
        self.hyper = torch.zeros(4, device=dev, dtype=torch.float32)
        self.hyper_host = torch.zeros(4, dtype=torch.float32).pin_memory() if dev.type == 'cuda' else torch.zeros(4)

    @property
    def t(self):
        return self.ts[0]

    @t.setter
    def t(self, v):
        self.ts = [v, v]

    def zero_grad(self):
        for s in self.segments:
            s.grad.zero_()

    def loss(self, restored, logits, labels, clean):
        n = len(logits)
        ce = sum(F.cross_entropy(logits[i], labels[i]) for i in range(n)) / n        # train.py:88
        l1 = l1_loss(restored, clean)                                                # train.py:89
        if self.decompose is not None:                                               # train.py:90-91
            l1 = l1 + self.w_freq * spectral_l1_loss(restored, clean, self.decompose)
        return l1 + self.w * ce, l1, ce                                              # train.py:92

    def _active(self):
        return self.segments[:1] if self.encoder_only else self.segments

    def _body(self, x_query, x_key, clean, use_hyper):
        self.zero_grad()
        if self.encoder_only:                                                        # train.py:84-87
            _, logits, labels, _ = self.net.E(x_query, x_key)
            n = len(logits)
            loss = ce = sum(F.cross_entropy(logits[i], labels[i]) for i in range(n)) / n
            l1 = torch.zeros((), device=ce.device)
        else:
            restored, logits, labels = self.net(x_query, x_key)
            loss, l1, ce = self.loss(restored, logits, labels, clean)
        loss.backward()
        if self.ddp is not None:
            self.ddp.finish()
        gscale = 1.0 / self.ddp.world if self.ddp is not None else 1.0       # mean over ranks, fused into Adam
        for i, s in enumerate(self._active()):
            if use_hyper:
                ops.adam_step_dev(s.flat, s.grad, s.m, s.v, self.hyper[2 * i:2 * i + 2], self.betas[0], self.betas[1],
                                  self.eps, gscale)
            else:
                ops.adam_step(s.flat, s.grad, s.m, s.v, self.lr, self.betas[0], self.betas[1], self.eps, self.ts[i], gscale)
        return dict(loss=loss.detach(), l1=l1.detach(), ce=ce.detach())

    def _tick(self, d=1):
        for i in range(len(self._active())):
            self.ts[i] += d

    def _advance(self):
        """step counts += 1 and the step-dependent Adam scalars {lr/(1-b1^t), 1/sqrt(1-b2^t)} of each segment pushed to
        the device."""
        self._tick()
        for i, t in enumerate(self.ts):
            if t >= 1:
                self.hyper_host[2 * i] = self.lr / (1.0 - self.betas[0] ** t)
                self.hyper_host[2 * i + 1] = 1.0 / (1.0 - self.betas[1] ** t) ** 0.5
        self.hyper.copy_(self.hyper_host, non_blocking=True)

    def step(self, x_query, x_key, clean):
        """One optimisation step; returns the (device) loss tensor.  Replays the captured graph when there is one."""
        if self.graph is not None:
            for dst, src in zip(self.static_in, (x_query, x_key, clean)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            self._advance()
            self.graph.replay()
            self.last = self.static_out
            return self.last['loss']
        self._tick()
        self.last = self._body(x_query, x_key, clean, False)
        return self.last['loss']

    def capture(self, x_query, x_key, clean, warmup=2):
        """Capture the whole step (zero_grad, forward, losses, backward, gradient all-reduce, Adam, momentum and queue
        updates: ~7 500 kernel launches) into ONE CUDA graph; later ``step`` calls copy the crops into the static input
        buffers and replay it.  ``warmup`` eager steps run first on a side stream (allocator warm-up, one-time
        cudaFuncSetAttribute calls); they are real optimisation steps."""
        self.static_in = [t.clone() for t in (x_query, x_key, clean)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._advance()
                self._body(*self.static_in, True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        self._advance()
        n0 = ops.launch_count()
        with torch.cuda.graph(graph):
            self.static_out = self._body(*self.static_in, True)
        self.graph_launches = ops.launch_count() - n0        # libfreqair kernels recorded in the graph (per replay)
        self._tick(-1)                  # capture records the step without executing it
        self.graph = graph
        self.last = self.static_out
        return self.static_out['loss']
