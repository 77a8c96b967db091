import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from helpers import make_args, load_params, digest
from oracle import cases, network as net
from rainbow_iqn_apex_b200 import Learner
from test_gpu_learn import FakeMem, _cfg, _dev_batch, _learner
dev = torch.device("cuda")
g = np.load(os.path.join(R, "tests/golden/iqn_small.npz"))
seed, batch, steps = int(g["seed"]), int(g["batch"]), int(g["steps"]); cfg = _cfg(g)
lr = _learner(dev, batch, cfg, net.make_params(seed))
for s in range(steps):
    b = cases.make_batch(seed + 10 + s, batch); taus = tuple(torch.from_numpy(t) for t in cases.make_taus(seed + 20 + s, batch, cfg))
    lr._inject = dict(noises=cases.make_noises(seed + 30 + s), taus=taus)
    st, ac, rt, nx, nt = _dev_batch(b, dev, fp32_frames=False)
    w = torch.from_numpy(b["weights"]).to(dev)
    _, loss = lr.learn(FakeMem((np.arange(batch), st, ac, rt, nx, nt, w)), None)
    print("step", s, "loss rel", np.max(np.abs(loss.cpu().numpy() - g[f"loss_{s}"]) / np.abs(g[f"loss_{s}"])))
    for k, p in lr.online_net.named_parameters():
        gd, ref = digest(p.grad), g[f"grad_{s}_{k}"]
        pd, pref = digest(p), g[f"param_{s}_{k}"]
        print(f"  {k:26s} grad l2 {gd[2]:.6e} ref {ref[2]:.6e}  rel {abs(gd[2]-ref[2])/ref[2]:.2e} | param head maxdiff {np.max(np.abs(pd[3:]-pref[3:])):.2e} l2 rel {abs(pd[2]-pref[2])/pref[2]:.2e}")
# checkpoint flat compare
lr.save("/tmp", 1, 2, "ck.pth")
args = make_args(dev, batch, cfg); args.model = "/tmp/ck.pth"
lr2 = Learner(args, 18, None)
d = (lr2.online_net._flat - lr.online_net._flat).abs()
print("flat maxdiff", float(d.max()), "n diff", int((d > 0).sum()), "nan", int(torch.isnan(lr.online_net._flat).sum()), int(torch.isnan(lr2.online_net._flat).sum()))
if (d > 0).any():
    idx = torch.nonzero(d > 0)[:5, 0]
    print("first idx", idx.tolist(), {k: p._riqn_offset for k, p in lr.online_net.named_parameters()})
