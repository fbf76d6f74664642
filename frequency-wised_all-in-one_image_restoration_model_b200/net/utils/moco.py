"""MoCo wrapper with the reference's interface and buffers (net/utils/moco.py:6-170): query / key encoders,
momentum update m=0.999, L queues [L, dim, K], logits / labels, enqueue.

What changes underneath: the key encoder's momentum update (99.7 M parameters streamed every step,
moco.py:45-50 - 489 tiny kernels in the reference) is ONE fused kernel over two flat parameter buffers that
the q / k parameters are views of.  The contrastive logits ([B, 1+K] per band) are a few KFLOP and stay on
the host side of the ABI.  The dead DDP helpers of the reference (moco.py:69-113,175-185; never called) are
not reproduced.
"""
import torch
import torch.nn as nn

from ... import ops


def flatten_parameters(params):
    """Re-home ``params`` (list of nn.Parameter) as views of one flat fp32 buffer; returns the buffer.
    Values are preserved; 16-byte alignment of every view is kept by padding to 4 floats."""
    params = list(params)
    if not params:
        return None
    sizes = [(p.numel() + 3) // 4 * 4 for p in params]
    flat = torch.empty(sum(sizes), device=params[0].device, dtype=torch.float32)
    flat.zero_()
    off = 0
    for p, n in zip(params, sizes):
        view = flat[off:off + p.numel()].view(p.shape)
        view.copy_(p.data)
        p.data = view
        off += n
    return flat


class MoCo(nn.Module):
    def __init__(self, opt, base_encoder, dim, K=3 * 256, m=0.999, T=0.07, mlp=False):
        super().__init__()
        self.num_losses = opt.L
        self.opt = opt
        self.K, self.m, self.T = K, m, T
        self.encoder_q = base_encoder(opt)
        self.encoder_k = base_encoder(opt)
        for param_q, param_k in zip(self.encoder_q.parameters(), self.encoder_k.parameters()):
            param_k.data.copy_(param_q.data)
            param_k.requires_grad = False
        self.register_buffer('queue', torch.randn(self.num_losses, dim, K))
        for i in range(self.num_losses):
            self.queue[i] = nn.functional.normalize(self.queue[i], dim=0)
        self.register_buffer('queue_ptr', torch.zeros(1, dtype=torch.long))
        self._flat_q = self._flat_k = None
        self._flat_k_rn = None          # TF32-rounded copy of the key encoder's weights (see rounded_key_weights)
        # the key encoder (no gradient) is independent of the query encoder: run it on a side stream so that its
        # latency-bound kernels (small grids at the 16x16 / 8x8 levels) fill the gaps of the query branch
        self.key_stream = None
        self.overlap_key_encoder = True

    def _ensure_flat(self):
        pq = list(self.encoder_q.parameters())
        ok = (self._flat_q is not None and self._flat_q.device == pq[0].device
              and pq[0].data_ptr() == self._flat_q.data_ptr())
        if not ok:
            self._flat_q = flatten_parameters(pq)
            self._flat_k = flatten_parameters(list(self.encoder_k.parameters()))
        return self._flat_q, self._flat_k

    def rounded_key_weights(self, owner):
        """Keep a TF32-rounded copy of the key encoder's flat buffer, refreshed right after every momentum update, and
        mark the key parameters with views of it (net/lewin.rn_weight): the 1xTF32 LeFF contractions of the key branch
        then read pre-rounded weights like those of the query branch and the restorer (trainer.TrainStep)."""
        _, fk = self._ensure_flat()
        self._flat_k_rn = torch.empty_like(fk)
        off = 0
        for p in self.encoder_k.parameters():
            p._fa_rn = self._flat_k_rn[off:off + p.numel()].view(p.shape)
            p._fa_rn_owner = owner
            off += (p.numel() + 3) // 4 * 4

    @torch.no_grad()
    def _momentum_update_key_encoder(self):
        fq, fk = self._ensure_flat()
        ops.momentum_update(fk, fq, self.m)
        if self._flat_k_rn is not None and self._flat_k_rn.numel() == fk.numel():
            ops.round_tf32(fk, self._flat_k_rn)

    @torch.no_grad()
    def _dequeue_and_enqueue(self, keys):
        """moco.py:53-66 without the reference's ``int(self.queue_ptr)`` device->host sync: the write columns are
        computed on the device from ``queue_ptr``, so the step never blocks the host and can be captured in a CUDA graph."""
        batch_size = keys[0].shape[0]
        assert self.K % batch_size == 0
        cols = (self.queue_ptr + torch.arange(batch_size, device=self.queue_ptr.device)) % self.K
        for i in range(len(keys)):
            self.queue[i].index_copy_(1, cols, keys[i].transpose(0, 1))
        self.queue_ptr.copy_((self.queue_ptr + batch_size) % self.K)

    def forward(self, im_q, im_k):
        if self.training:
            side = None
            if self.overlap_key_encoder and im_q.is_cuda:
                if self.key_stream is None:
                    self.key_stream = torch.cuda.Stream()
                side = self.key_stream
                side.wait_stream(torch.cuda.current_stream())
            with torch.no_grad(), torch.cuda.stream(side):
                self._momentum_update_key_encoder()
                _, k, _ = self.encoder_k(im_k)
                k = [nn.functional.normalize(ki, dim=1) for ki in k]
            embedding, q, inter = self.encoder_q(im_q)
            n = min(self.num_losses, len(q))
            q = [nn.functional.normalize(q[i], dim=1) for i in range(n)]
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)
                for ki in k:
                    ki.record_stream(torch.cuda.current_stream())
            k = k[:n]
            logits, labels = [], []
            for i in range(n):
                l_pos = torch.einsum('nc,nc->n', [q[i], k[i]]).unsqueeze(-1)
                l_neg = torch.einsum('nc,ck->nk', [q[i], self.queue[i].clone().detach()])
                lg = torch.cat([l_pos, l_neg], dim=1) / self.T
                logits.append(lg)
                labels.append(torch.zeros(lg.shape[0], dtype=torch.long, device=lg.device))
            self._dequeue_and_enqueue(k)
            return embedding, logits, labels, inter
        # eval: the reference computes and discards the contrastive heads (moco.py:167-170); encoders that
        # expose `trunk` skip them (11.3 GFLOP + 805 MB of activations per 16 tiles)
        if hasattr(self.encoder_q, 'trunk') and getattr(self.encoder_q, 'embedding_is_none', True):
            return None, self.encoder_q.trunk(im_q)
        embedding, _, inter = self.encoder_q(im_q)
        return embedding, inter
