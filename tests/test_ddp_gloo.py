"""CPU, world_size 2, gloo: the host logic of the data-parallel step - flat parameter segments, bucket partition,
post-accumulate-grad hooks, asynchronous launches, finish() - without a GPU.  The CUDA path only adds a side stream."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from conftest import PKG_NAME


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


class _Reversed(nn.Module):
    """Parameters REGISTERED as [out, inp] but USED inp -> out: backward completes `out`'s gradient first, so with small
    buckets a parameter that straddles a bucket boundary has its head in a bucket whose other parameters finish earlier.
    (The round-1 bookkeeping counted a parameter only in the bucket of its last element and reduced such a bucket
    before the straddling gradient was written - found by the advisor; nn.Sequential's forward-order registration hid it.)"""

    def __init__(self):
        super().__init__()
        self.out = nn.Linear(5, 3)
        self.mid = nn.Linear(13, 5)
        self.inp = nn.Linear(7, 13)

    def forward(self, x):
        return self.out(torch.tanh(self.mid(torch.tanh(self.inp(x)))))


def _model(variant='seq'):
    torch.manual_seed(0)
    if variant == 'reversed':
        return _Reversed()
    return nn.Sequential(nn.Linear(7, 13), nn.Tanh(), nn.Linear(13, 5), nn.Tanh(), nn.Linear(5, 3))


def _worker(rank, world, port, out, variant='seq', bucket_floats=16):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        trainer = importlib.import_module(PKG_NAME + '.trainer')
        net = _model(variant)
        params = list(net.parameters())
        seg_a, seg_b = trainer.Segment(params[:2]), trainer.Segment(params[2:])
        # 64-byte (or 96-byte) buckets: several buckets per segment, parameters straddling bucket boundaries
        ddp = trainer.BucketedAllReduce([seg_a, seg_b], bucket_mb=4 * bucket_floats / (1 << 20))
        assert len(ddp.buckets) > 4
        res = []
        for step in range(2):                          # two steps: reset() must re-arm the hooks
            seg_a.grad.zero_(); seg_b.grad.zero_()
            x = torch.randn(4, 7, generator=torch.Generator().manual_seed(10 * step + rank))
            net(x).square().sum().backward()
            ddp.finish()
            res.append(torch.cat([seg_a.grad, seg_b.grad]).clone())
        # parameters are still views of the flat buffers and their .grad views of the flat gradients
        assert params[0].data_ptr() == seg_a.flat.data_ptr() and params[0].grad.data_ptr() == seg_a.grad.data_ptr()
        out[rank] = torch.stack(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('variant,bucket_floats', [('seq', 16), ('reversed', 24), ('reversed', 16)])
def test_bucketed_allreduce_world2(variant, bucket_floats):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, variant, bucket_floats), nprocs=world, join=True)
    assert torch.equal(out[0], out[1])                  # both ranks hold the same reduced gradients
    # expected: sum over ranks of the single-process gradients, in flat-segment order
    trainer = importlib.import_module(PKG_NAME + '.trainer')
    for step in range(2):
        tot = None
        for rank in range(world):
            net = _model(variant)
            x = torch.randn(4, 7, generator=torch.Generator().manual_seed(10 * step + rank))
            net(x).square().sum().backward()
            params = list(net.parameters())
            flat = []
            for group in (params[:2], params[2:]):
                offs, total = trainer._offsets(group)
                buf = torch.zeros(total)
                for p, o in zip(group, offs):
                    buf[o:o + p.numel()] = p.grad.flatten()
                flat.append(buf)
            g = torch.cat(flat)
            tot = g if tot is None else tot + g
        assert torch.allclose(out[0][step], tot, rtol=1e-6, atol=1e-7)
