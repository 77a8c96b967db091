import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from helpers import make_args, load_params
from oracle import cases, losses, network as net
from rainbow_iqn_apex_b200 import Learner, compute_loss_iqn
dev = torch.device("cuda")
B, cfg, seed = 4, cases.iqn_cfg(8, 8, 4), 101
params = net.make_params(seed)
p_on, p_tg = net.to_torch(params, requires_grad=True), net.to_torch(params)
adam = losses.Adam([k for k in p_on if net.is_trainable(k)], lr=5e-5, eps=3.125e-4)
def data(s):
    b = cases.make_batch(seed + 10 + s, B); taus = tuple(torch.from_numpy(t) for t in cases.make_taus(seed + 20 + s, B, cfg))
    return b, taus, cases.make_noises(seed + 30 + s)
b, taus, noises = data(0)
losses.learn_step(p_on, p_tg, adam, cases.batch_to_torch(b), torch.from_numpy(b["weights"]), noises, taus, cfg)
params1 = {k: v.detach().numpy().copy() for k, v in p_on.items()}
g = np.load(os.path.join(R, "tests/golden/iqn_small.npz"))
for mode in ("fresh_with_step1_params_target_initial", "fresh_with_step1_params_target_synced"):
    lr = Learner(make_args(dev, B, cfg), 18, None)
    load_params(lr.online_net, params1)
    if mode.endswith("initial"):
        load_params(lr.target_net, params)
    else:
        lr.update_target_net()
    b, taus, noises = data(1)
    lr._inject = dict(noises=noises, taus=taus)
    st, nx = torch.from_numpy(b["states"]).to(dev), torch.from_numpy(b["next_states"]).to(dev)
    ac, rt, nt = (torch.from_numpy(b[k]).to(dev) for k in ("actions", "returns", "nonterminals"))
    w = torch.from_numpy(b["weights"]).to(dev)
    loss, dtheta, keep, _ = compute_loss_iqn.loss_core(lr, st, ac, rt, nx, nt, keep_graph=True)
    lr.online_net.zero_grad(); lr.online_net.backward_iqn(keep, dtheta, w / B, ac)
    print(mode, "loss rel vs golden step1", np.max(np.abs(loss.cpu().numpy() - g["loss_1"]) / np.abs(g["loss_1"])))
    for k in ("conv1.weight", "conv2.weight", "conv3.weight", "iqn_fc.weight"):
        gg = dict(lr.online_net.named_parameters())[k].grad
        l2 = float(gg.double().norm()); ref = g[f"grad_1_{k}"][2]
        print(f"   {k:16s} l2 {l2:.6e} golden {ref:.6e} rel {abs(l2-ref)/ref:.2e}")
# oracle step 1 directly
b, taus, noises = data(1)
o_loss, o_grads = losses.learn_step(p_on, p_tg, adam, cases.batch_to_torch(b), torch.from_numpy(b["weights"]), noises, taus, cfg)
for k in ("conv1.weight", "conv2.weight", "conv3.weight"):
    print("oracle step1", k, float(o_grads[k].double().norm()), "golden", g[f"grad_1_{k}"][2])
