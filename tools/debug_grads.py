"""Dev aid: per-parameter gradient / intermediate comparison of the CUDA learner step vs the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import make_args, load_params
from oracle import cases, losses, network as net
from rainbow_iqn_apex_b200 import Learner, compute_loss_iqn

dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = cases.iqn_cfg(8, 8, 4)
seed = 101
params = net.make_params(seed)
lr = Learner(make_args(dev, B, cfg), 18, None)
load_params(lr.online_net, params); lr.update_target_net()
b = cases.make_batch(seed + 10, B); taus = tuple(torch.from_numpy(t) for t in cases.make_taus(seed + 20, B, cfg))
noises = cases.make_noises(seed + 30)
lr._inject = dict(noises=noises, taus=taus)
st, nx = torch.from_numpy(b["states"]).to(dev), torch.from_numpy(b["next_states"]).to(dev)
ac, rt, nt = (torch.from_numpy(b[k]).to(dev) for k in ("actions", "returns", "nonterminals"))
w = torch.from_numpy(b["weights"]).to(dev)
dbg = {}
loss, dtheta, keep, _ = compute_loss_iqn.loss_core(lr, st, ac, rt, nx, nt, keep_graph=True, debug=dbg)
lr.online_net.zero_grad()
lr.online_net.backward_iqn(keep, dtheta, w / B, ac)
torch.cuda.synchronize()
p_on, p_tg = net.to_torch(params, requires_grad=True), net.to_torch(params)
ok = {}
o_loss = losses.iqn_loss(p_on, p_tg, *cases.batch_to_torch(b), noises, taus, **cfg, keep=ok)
for k in ("feat", "x", "h_v", "h_a", "q"):
    ok[k].retain_grad()
ok["theta"].retain_grad()
(torch.from_numpy(b["weights"]) * o_loss).mean().backward()
print("loss relerr", float(((loss.cpu() - o_loss.detach()).abs() / o_loss.detach().abs()).max()))
N = cfg["n_tau"]
dth_ref = ok["theta"].grad.t().reshape(-1)          # (N,B) row = i*B+b
print("dtheta*g relerr", float(((dtheta.cpu() * (w.cpu() / B).repeat(N)) - dth_ref).abs().max() / dth_ref.abs().max()))
for k, p in lr.online_net.named_parameters():
    g, r = p.grad.cpu(), p_on[k].grad
    print(f"{k:28s} relnorm {float((g - r).norm() / (r.norm() + 1e-30)):.3e}  |ref| {float(r.norm()):.3e}")
