import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from helpers import make_args, load_params
from oracle import cases, losses, network as net
from rainbow_iqn_apex_b200 import Learner, compute_loss_iqn
from rainbow_iqn_apex_b200.model import DQN
dev = torch.device("cuda")
B, cfg, seed = 4, cases.iqn_cfg(8, 8, 4), 101
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
params = net.make_params(seed)
params0 = params
if len(sys.argv) > 2:   # use oracle's post-step-0 params for the online net
    p_on0, p_tg0 = net.to_torch(params, requires_grad=True), net.to_torch(params)
    adam0 = losses.Adam([k for k in p_on0 if net.is_trainable(k)], lr=5e-5, eps=3.125e-4)
    b0 = cases.make_batch(seed + 10, B); taus0 = tuple(torch.from_numpy(t) for t in cases.make_taus(seed + 20, B, cfg))
    losses.learn_step(p_on0, p_tg0, adam0, cases.batch_to_torch(b0), torch.from_numpy(b0["weights"]), cases.make_noises(seed + 30), taus0, cfg)
    params = {k: v.detach().numpy().copy() for k, v in p_on0.items()}
b = cases.make_batch(seed + 10 + S, B); taus = tuple(torch.from_numpy(t) for t in cases.make_taus(seed + 20 + S, B, cfg)); noises = cases.make_noises(seed + 30 + S)
lr = Learner(make_args(dev, B, cfg), 18, None); load_params(lr.online_net, params); load_params(lr.target_net, params0)
lr._inject = dict(noises=noises, taus=taus)
stash = {}
orig = DQN.backward_trunk
def patched(self, keep, dfeat):
    stash["dfeat"] = dfeat.clone(); stash["keep"] = keep
    return orig(self, keep, dfeat)
DQN.backward_trunk = patched
st, nx = torch.from_numpy(b["states"]).to(dev), torch.from_numpy(b["next_states"]).to(dev)
ac, rt, nt = (torch.from_numpy(b[k]).to(dev) for k in ("actions", "returns", "nonterminals"))
w = torch.from_numpy(b["weights"]).to(dev)
loss, dtheta, keep, _ = compute_loss_iqn.loss_core(lr, st, ac, rt, nx, nt, keep_graph=True)
lr.online_net.zero_grad(); lr.online_net.backward_iqn(keep, dtheta, w / B, ac)
p_on, p_tg = net.to_torch(params, requires_grad=True), net.to_torch(params0)
ok = {}
o_loss = losses.iqn_loss(p_on, p_tg, *cases.batch_to_torch(b), noises, taus, **cfg, keep=ok)
for k in ("o1", "o2", "o3", "feat", "x"): ok[k].retain_grad()
(torch.from_numpy(b["weights"]) * o_loss).mean().backward()
def cmp(name, a, r):
    a, r = a.detach().cpu().double().reshape(-1), r.detach().double().reshape(-1)
    d = (a - r).abs()
    print(f"{name:10s} relnorm {float((a-r).norm()/r.norm()):.3e} maxabs {float(d.max()):.3e} at {int(d.argmax())} got {float(a[d.argmax()]):.4e} ref {float(r[d.argmax()]):.4e}  nbad {(d > 1e-3*r.abs().max()).sum().item()}")
(g1, g2, g3), (col1, col2, col3), (out1, out2, out3) = keep["g"], keep["col"], keep["out"]
cmp("out1", out1, ok["o1"]); cmp("out2", out2, ok["o2"]); cmp("out3", out3, ok["o3"])
cmp("dfeat", stash["dfeat"], ok["feat"].grad * (ok["feat"] > 0))
cmp("dfeat_raw", stash["dfeat"], ok["feat"].grad)
from rainbow_iqn_apex_b200._lib import call, ptr
d = lr.online_net; gv = d.grad_view
d_out2 = torch.empty_like(out2); dY3 = torch.empty(B * 49, 64, device=dev); dcol3 = torch.empty_like(col3)
tmpw = torch.zeros_like(d.conv3.weight); tmpb = torch.zeros_like(d.conv3.bias)
call("riqn_conv_bwd", g3, ptr(stash["dfeat"]), ptr(out3), ptr(col3), ptr(d.conv3.weight), ptr(dY3), ptr(dcol3), ptr(tmpw), ptr(tmpb), ptr(d_out2))
cmp("d_out2", d_out2, ok["o2"].grad)
cmp("d_out2*m", d_out2 * (out2 > 0), ok["o2"].grad * (ok["o2"] > 0))
print("mask mismatches out2", int(((out2.cpu() > 0) != (ok["o2"] > 0)).sum()), "out1", int(((out1.cpu() > 0) != (ok["o1"] > 0)).sum()), "out3", int(((out3.cpu() > 0) != (ok["o3"] > 0)).sum()))

for k in ("conv1.weight", "conv2.weight", "conv2.bias", "conv3.weight"):
    cmp(k, dict(d.named_parameters())[k].grad, p_on[k].grad)
dY2 = torch.empty(B * 81, 64, device=dev); dcol2 = torch.empty_like(col2); d_out1 = torch.empty_like(out1)
tw = torch.zeros_like(d.conv2.weight); tb = torch.zeros_like(d.conv2.bias)
call("riqn_conv_bwd", g2, ptr(d_out2), ptr(out2), ptr(col2), ptr(d.conv2.weight), ptr(dY2), ptr(dcol2), ptr(tw), ptr(tb), ptr(d_out1))
cmp("conv2.w(re)", tw, p_on["conv2.weight"].grad); cmp("conv2.b(re)", tb, p_on["conv2.bias"].grad)
cmp("d_out1", d_out1, ok["o1"].grad)
