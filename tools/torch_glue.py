"""Which torch (non-libfreqair) kernels remain in one train step: torch.profiler CUDA kernel table."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
PKG = bench.PKG
synth = importlib.import_module(PKG + '.synth'); model = importlib.import_module(PKG + '.net.model'); trainer = importlib.import_module(PKG + '.trainer')
torch.manual_seed(0)
net = model.AirNet(bench.make_opt(16)).cuda().train()
ts = trainer.TrainStep(net, lr=2e-4, contrast_loss_weight=0.6)
x = [t.cuda() for t in synth.noisy_batch(16, 25)]
for _ in range(3): ts.step(*x)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ts.step(*x); torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows); n = sum(r[1] for r in rows)
print(f'total CUDA kernel time {tot:.1f} ms in {n} launches')
ours = [r for r in rows if 'anonymous' in r[0] or 'unnamed' in r[0]]
print(f'libfreqair: {sum(r[2] for r in ours):.1f} ms in {sum(r[1] for r in ours)} launches')
print('--- torch / library kernels')
for k, c, ms in [r for r in rows if r not in ours][:30]:
    print(f'{ms:8.3f} ms {c:5d}  {k[:110]}')
