import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, torch.nn.functional as F
from helpers import make_args, load_params
from oracle import cases, network as net
from rainbow_iqn_apex_b200.model import DQN
from rainbow_iqn_apex_b200._lib import call, ptr
dev = torch.device("cuda")
B = 4
params = net.make_params(101)
d = DQN(make_args(dev, B), 18).to(dev); load_params(d, params)
b = cases.make_batch(111, B)
x = torch.from_numpy(b["states"]).to(dev)
p = net.to_torch(params, requires_grad=True)
xin = torch.from_numpy(b["states"]).float().div_(255)
o1 = F.relu(F.conv2d(xin, p["conv1.weight"], p["conv1.bias"], stride=4, padding=1)); o1.retain_grad()
o2 = F.relu(F.conv2d(o1, p["conv2.weight"], p["conv2.bias"], stride=2)); o2.retain_grad()
o3 = F.relu(F.conv2d(o2, p["conv3.weight"], p["conv3.bias"]))
gfeat = torch.from_numpy(np.random.RandomState(0).standard_normal((B, 3136)).astype(np.float32))
(o3.reshape(B, -1) * gfeat).sum().backward()
for trial in range(3):
    junk = torch.randn(50_000_000, device=dev)  # dirty the allocator pool
    del junk
    keep = {}
    feat = d.trunk(x, keep)
    print("trial", trial, "feat err", float((feat.cpu() - o3.detach().reshape(B, -1)).abs().max()))
    d.zero_grad()
    # replicate backward_trunk with intermediates exposed
    (g1, g2, g3), (col1, col2, col3), (out1, out2, out3) = keep["g"], keep["col"], keep["out"]
    gv = d.grad_view
    dfeat = gfeat.to(dev)
    d_out2 = torch.full_like(out2, float("nan")); dY3 = torch.full((B * 49, 64), float("nan"), device=dev); dcol3 = torch.full_like(col3, float("nan"))
    call("riqn_conv_bwd", g3, ptr(dfeat), ptr(out3), ptr(col3), ptr(d.conv3.weight), ptr(dY3), ptr(dcol3), ptr(gv(d.conv3.weight)), ptr(gv(d.conv3.bias)), ptr(d_out2))
    ref_d2 = o2.grad
    print("   d_out2 err", float((d_out2.cpu() - ref_d2).abs().max()), "nan", int(torch.isnan(d_out2).sum()), "dcol3 nan", int(torch.isnan(dcol3).sum()), "dY3 nan", int(torch.isnan(dY3).sum()))
    d_out1 = torch.full_like(out1, float("nan")); dY2 = torch.full((B * 81, 64), float("nan"), device=dev); dcol2 = torch.full_like(col2, float("nan"))
    call("riqn_conv_bwd", g2, ptr(d_out2), ptr(out2), ptr(col2), ptr(d.conv2.weight), ptr(dY2), ptr(dcol2), ptr(gv(d.conv2.weight)), ptr(gv(d.conv2.bias)), ptr(d_out1))
    print("   d_out1 err", float((d_out1.cpu() - o1.grad).abs().max()), "nan", int(torch.isnan(d_out1).sum()))
    print("   conv2.w grad err", float((gv(d.conv2.weight).cpu() - p["conv2.weight"].grad).abs().max()), "conv3.w", float((gv(d.conv3.weight).cpu() - p["conv3.weight"].grad).abs().max()))
