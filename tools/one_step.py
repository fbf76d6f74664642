"""One configs[1] train step launched kernel by kernel (no CUDA graph) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum` launch lists (profiles/).  Warm-up steps run unprofiled."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
PKG = bench.PKG
synth = importlib.import_module(PKG + '.synth'); model = importlib.import_module(PKG + '.net.model'); trainer = importlib.import_module(PKG + '.trainer')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
net = model.AirNet(bench.make_opt(B)).cuda().train()
ts = trainer.TrainStep(net, lr=2e-4, contrast_loss_weight=0.6)
x = [t.cuda() for t in synth.noisy_batch(B, 25)]
for _ in range(2):
    ts.step(*x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = ts.step(*x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('loss', float(loss))
