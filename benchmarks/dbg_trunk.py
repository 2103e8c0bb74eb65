import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_trunks
from megreader_b200 import conv_engine
cfg = int(sys.argv[1])
cuda = torch.device("cuda:0")
net, _ = bench_trunks.build(cfg, cuda, engine=False)
for m in net.modules():
    if isinstance(m, torch.nn.BatchNorm2d): m.eval()
x, y, l = [t.to(cuda) for t in bench_trunks.synth(1, 4, bench_trunks.CFG[cfg]["hw"], 8)]
state = {k: v.clone() for k, v in net.state_dict().items()}
def run():
    for p in net.parameters(): p.grad = None
    torch.manual_seed(1); np.random.seed(1)
    loss, _ = net(x, y, l); loss = loss.mean(); loss.backward()
    g = {n: p.grad.detach().float().clone() for n, p in net.named_parameters() if p.grad is not None}
    net.load_state_dict(state)
    return float(loss), g
l0, g0 = run()
l0b, g0b = run()
try:
    conv_engine.use_engine_convs(net)
    l1, g1 = run()
except Exception as e:
    import traceback; traceback.print_exc(); sys.exit(0)
print("loss", l0, l0b, l1)
for n in g0:
    r = g0[n]
    if r.numel() >= 4096:
        c = float((r * g1[n]).sum() / (r.norm() * g1[n].norm() + 1e-20))
        c0 = float((r * g0b[n]).sum() / (r.norm() * g0b[n].norm() + 1e-20))
        print("%-50s cos %.4f (lib-lib %.4f) norm %.3e %.3e" % (n, c, c0, float(r.norm()), float(g1[n].norm())))
