"""One CRNN training step at the bench shape inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off).
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X \
        python benchmarks/crnn_step_profile.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.BATCH_PER_GPU
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
net = bench.build_model(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
x, y, l = [t.to(dev) for t in bench.synth_batch(0, n)]


from megreader_b200 import crnn_engine  # noqa: E402
crnn_engine.set_compute_dtype(torch.bfloat16)


def step():
    opt.zero_grad(set_to_none=True)
    loss, _ = net(x, y, l)
    loss.mean().backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loss = step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss.mean()))
