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
        self.flat_rn = None             # TF32-rounded copy of `flat` (TrainStep: weight operand of the 1xTF32 contractions)
        for p, o in zip(self.params, offs):
            p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def make_rounded_copy(self, owner):
        self.flat_rn = torch.empty_like(self.flat)
        for p, o in zip(self.params, self.offs):
            p._fa_rn = self.flat_rn[o:o + p.numel()].view(p.shape)
            p._fa_rn_owner = owner


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
        self.launched = []
        self.late = []
        per = int(bucket_mb * (1 << 20) // 4)
        for seg in segments:
            n = seg.grad.numel()
            nb = max(1, (n + per - 1) // per)
            bounds = [min(n, i * per) for i in range(nb + 1)]
            base = len(self.buckets)
            counts = [0] * nb
            owners = []
            for p, o in zip(seg.params, seg.offs):
                # a parameter that straddles a bucket boundary has elements in BOTH slices: neither may be reduced
                # before its gradient is complete, so it counts in every bucket it overlaps
                b0, b1 = min(nb - 1, o // per), min(nb - 1, (o + max(p.numel(), 1) - 1) // per)
                for b in range(b0, b1 + 1):
                    counts[b] += 1
                owners.append(tuple(range(base + b0, base + b1 + 1)))
            for i in range(nb):
                self.buckets.append((seg.grad[bounds[i]:bounds[i + 1]], counts[i]))
            for p, bs in zip(seg.params, owners):
                p.register_post_accumulate_grad_hook(self._make_hook(bs, id(p)))
                self.bucket_of[id(p)] = bs
        self.reset()

    def param_ready(self, p):
        """A block-level backward accumulated this parameter's gradient straight into the flat buffer (no
        AccumulateGrad, so no hook): same bookkeeping as the hook.  Exactly once per parameter per step: a weight
        shared by two direct-sink nodes would be announced by the first while the second still has to add to it."""
        k = id(p)
        bs = self.bucket_of.get(k)
        if bs is None:
            return
        if k in self.seen:
            raise RuntimeError('BucketedAllReduce: a parameter was reported complete twice in one step (a weight shared '
                               'by several direct-gradient nodes is not supported under data parallelism)')
        self.seen.add(k)
        self._dec(bs)          # (the AccumulateGrad hook that follows for the same parameter is dropped, see _make_hook)

    def _dec(self, bs):
        for b in bs:
            if self.launched[b]:
                # a gradient reported complete AFTER its bucket was reduced: the bookkeeping counted something twice
                self.late.append(b)
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._launch(b)

    def reset(self):
        self.seen = set()
        self.pending = [c for _, c in self.buckets]
        self.handles = []
        self.launched = [False] * len(self.buckets)
        self.late = []

    def _make_hook(self, bs, key):
        def hook(_param):
            # The engine runs AccumulateGrad - and with it this hook - for a parameter even when the block-level backward
            # returned None for it (its gradient went straight into the flat buffer and was announced by param_ready).
            # Hook and direct report therefore share one "seen" set: whichever comes first counts, the other is dropped.
            # (Round 1 counted both; buckets reached zero early and were reduced before their last gradients landed -
            # found by `bench.py --check-grads` on 2 GPUs.)
            if key in self.seen:
                return
            self.seen.add(key)
            self._dec(bs)
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
        if self.late:
            late, self.late = self.late, []
            raise RuntimeError(f'BucketedAllReduce: gradients were reported complete after their bucket had been reduced '
                               f'(buckets {sorted(set(late))}): a parameter was counted twice')
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
        self.ts = [0, 0]                # Adam step count of the encoder / decoder segment (host mirror)
        moco = net.E.E
        fq, _ = moco._ensure_flat()
        enc_params = [p for p in moco.encoder_q.parameters()]
        dec_params = [p for p in net.R.parameters()]
        self.segments = [Segment(enc_params, fq), Segment(dec_params)]
        moco._flat_q = self.segments[0].flat
        self.ddp = BucketedAllReduce(self.segments, bucket_mb) if distributed else None
        # Parameters now own .grad views of per-step-zeroed flat buffers: let the block backwards accumulate into them.
        # The switch is a per-PARAMETER attribute (net/lewin._sink), so a second TrainStep, or a plain optim.Adam loop
        # on another net in the same process, is not rewired by this constructor.
        ready = self.ddp.param_ready if self.ddp is not None else None
        for seg in self.segments:
            for p in seg.params:
                p._fa_direct = True
                p._fa_ready = ready
        # one TF32-rounded copy of each flat parameter buffer, refreshed at the start of every step (net/lewin.rn_weight)
        if self.segments[0].flat.is_cuda and any(b in (ops.GEMM_1X, ops.GEMM_2X) for b in (ops.LEFF_BACKEND, ops.LEFF_ENC_BACKEND)):
            for seg in self.segments:
                seg.make_rounded_copy(id(self))
            moco.rounded_key_weights(id(self))
        self.last = {}
        # CUDA-graph state (capture()): static inputs / loss
        self.graph = None
        self.graph_launches = 0
        self.graph_cfg = None
        self.static_in = None
        self.static_out = None
        dev = self.segments[0].flat.device
        # Optimiser state that changes from step to step lives ON THE DEVICE: per segment {lr (fp32), step count (int32)}.
        # fa_adam_tick increments the count and fa_adam_step_state derives the bias corrections from it inside the
        # kernel, so neither the eager step nor a graph replay reads host memory that a later step may already have
        # overwritten (the host runs many replays ahead of the GPU).
        self.opt_state = torch.zeros(2 * len(self.segments), device=dev, dtype=torch.float32)
        self.set_lr(lr)

    def set_lr(self, lr):
        """Learning rate of every segment (train.py's StepLR): a stream-ordered device write, valid under graph replay."""
        self.lr = lr
        self.opt_state[0::2].fill_(lr)

    def _state(self, i):
        return self.opt_state[2 * i:2 * i + 2]

    def _sync_counts(self):
        """Push the host mirror of the step counts to the device (after `ts.t = n`, checkpoint restore)."""
        cnt = self.opt_state.view(torch.int32)
        for i, t in enumerate(self.ts):
            cnt[2 * i + 1].fill_(int(t))

    @property
    def t(self):
        return self.ts[0]

    @t.setter
    def t(self, v):
        self.ts = [v, v]
        self._sync_counts()

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

    def _body(self, x_query, x_key, clean):
        from .net import lewin
        lewin.RN_OWNER[0] = id(self)
        try:
            return self._body_inner(x_query, x_key, clean)
        finally:
            lewin.RN_OWNER[0] = None

    def _body_inner(self, x_query, x_key, clean):
        self.zero_grad()
        for s in self._active():
            if s.flat_rn is not None:
                ops.round_tf32(s.flat, s.flat_rn)
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
            ops.adam_tick(self._state(i))
            ops.adam_step_state(s.flat, s.grad, s.m, s.v, self._state(i), self.betas[0], self.betas[1], self.eps, gscale)
        return dict(loss=loss.detach(), l1=l1.detach(), ce=ce.detach())

    def _tick(self, d=1):
        for i in range(len(self._active())):
            self.ts[i] += d

    def _cfg(self, x_query):
        return (self.encoder_only, tuple(x_query.shape), self.net.training, self.decompose is not None)

    def step(self, x_query, x_key, clean):
        """One optimisation step; returns the (device) loss tensor.  Replays the captured graph when there is one."""
        if self.graph is not None:
            if self._cfg(x_query) != self.graph_cfg:
                raise RuntimeError(f'TrainStep.step: the captured graph was recorded for (encoder_only, input shape, '
                                   f'training, spectral_l1) = {self.graph_cfg} but the step is now {self._cfg(x_query)}; '
                                   'call capture() again (e.g. at the epochs_encoder phase switch, train.py:82)')
            for dst, src in zip(self.static_in, (x_query, x_key, clean)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            self._tick()
            self.graph.replay()
            self.last = self.static_out
            return self.last['loss']
        self._tick()
        self.last = self._body(x_query, x_key, clean)
        return self.last['loss']

    def capture(self, x_query, x_key, clean, warmup=2):
        """Capture the whole step (zero_grad, forward, losses, backward, gradient all-reduce, step-count tick, Adam,
        momentum and queue updates) into ONE CUDA graph; later ``step`` calls copy the crops into the static input
        buffers and replay it.  ``warmup`` eager steps run first on a side stream (allocator warm-up, one-time
        cudaFuncSetAttribute calls); they are real optimisation steps.  The graph is valid for the configuration it was
        captured with (phase, shapes, train mode): ``step`` refuses to replay it under another one.

        Adam step counts: one count per SEGMENT (query encoder, restorer).  torch.optim.Adam keeps one per parameter
        and starts it at the first step in which the parameter receives a gradient; the two agree because every
        parameter of a segment receives a gradient in every step in which its segment is active (encoder-only phase:
        the restorer segment is skipped as a whole, exactly like parameters without .grad in torch)."""
        self.static_in = [t.clone() for t in (x_query, x_key, clean)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._tick()
                self._body(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(graph):
            self.static_out = self._body(*self.static_in)
        self.graph_launches = ops.launch_count() - n0        # libfreqair kernels recorded in the graph (per replay)
        self.graph = graph                                    # capture records the step without executing it
        self.graph_cfg = self._cfg(x_query)
        self.last = self.static_out
        return self.static_out['loss']
