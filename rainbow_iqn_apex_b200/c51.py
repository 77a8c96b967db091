"""Rainbow-only (C51) branch of the reference: DQN.forward (model.py:120-129) and the categorical loss
(agent.py:77-141), on the same CUDA trunk / NoisyLinear ops as the IQN path plus csrc/c51.cu."""
import torch

from ._lib import call, ptr
from .model import FEAT


def forward(net, x, log=False, keep=None, fresh_weights=False, want_argmax=None, support=None):
    """Returns probabilities (or log-probabilities) (B, A, atoms).  model.py:120-129"""
    if not fresh_weights:
        net.compose_weights()
    feat = net.trunk(x, keep)
    B = feat.shape[0]
    dev = feat.device
    hid, A, atoms = net.hidden, net.action_space, net.atoms
    h = torch.empty(B, 2 * hid, device=dev)
    call("riqn_noisy_linear_fwd", B, FEAT, 2 * hid, ptr(feat), ptr(net._w_eff_h), ptr(net._b_eff_h), ptr(h))
    zv = torch.empty(B, atoms, device=dev)
    za = torch.empty(B, A * atoms, device=dev)
    wz, bz = net._w_eff_z, net._b_eff_z          # rows [0, atoms) = z_v, rows [atoms, atoms + A*atoms) = z_a
    hv, ha = h[:, :hid], h[:, hid:]
    call("riqn_linear_fwd_ld", B, hid, atoms, ptr(hv), 2 * hid, ptr(wz), ptr(bz), ptr(zv), atoms, 0)
    wza, bza = wz[atoms:], bz[atoms:]
    call("riqn_linear_fwd_ld", B, hid, A * atoms, ptr(ha), 2 * hid, ptr(wza), ptr(bza), ptr(za), A * atoms, 0)
    out = torch.empty(B, A, atoms, device=dev)
    if support is None:
        support = net._support(dev)
    call("riqn_c51_head_fwd", B, A, atoms, ptr(zv), ptr(za), ptr(support), None if log else ptr(out),
         ptr(out) if log else None, ptr(want_argmax))
    if keep is not None:
        keep.update(feat=feat, h=h, B=B)
    return out


def loss_core(agent, states, actions, returns, next_states, nonterminals, debug=None):
    """agent.py:77-141.  Returns (loss (B,), backward(gscale) closure)."""
    from .compute_loss_iqn import _as_device_inputs
    states, actions, returns, next_states, nonterminals = _as_device_inputs(
        agent, states, actions, returns, next_states, nonterminals)
    on, tg = agent.online_net, agent.target_net
    B, A, atoms = states.shape[0], agent.action_space, agent.atoms
    dev = states.device
    inj = getattr(agent, "_inject", None)
    noises = inj["noises"] if inj else (None, None, None)
    on.reset_noise(noises[0])                                              # agent.py:82
    keep = {}
    log_ps = forward(on, states, log=True, keep=keep, fresh_weights=True, support=agent.support)   # :83
    # the gradient pass must see the weights of THIS noise sample: snapshot what backward needs
    w_h = on._w_eff_h.clone()
    w_z = on._w_eff_z.clone()
    eps = {n: (m.weight_epsilon.clone(), m.bias_epsilon.clone()) for n, m in on.noisy_layers()}
    on.reset_noise(noises[1])                                              # :95
    a_star = torch.empty(B, dtype=torch.int64, device=dev)
    forward(on, next_states, fresh_weights=True, want_argmax=a_star, support=agent.support)         # :97-102
    tg.reset_noise(noises[2])                                              # :103
    pns = forward(tg, next_states, fresh_weights=True, support=agent.support)                      # :104
    loss = torch.empty(B, device=dev)
    dq = torch.empty(B, atoms, device=dev)
    m_out = torch.empty(B, atoms, device=dev) if debug is not None else None
    call("riqn_c51_loss_fwd_bwd", B, A, atoms, ptr(log_ps), ptr(pns), ptr(actions), ptr(a_star), ptr(returns),
         ptr(nonterminals), ptr(agent.support), float(agent.discount ** agent.n), float(agent.Vmin), float(agent.Vmax),
         float(agent.delta_z), ptr(loss), ptr(dq), ptr(m_out))
    if debug is not None:
        debug.update(a_star=a_star, m=m_out, log_ps=log_ps)

    def backward(gscale):
        hid = on.hidden
        gv = on.grad_view
        hvL, haL, zvL, zaL = on.fcnoisy_h_v, on.fcnoisy_h_a, on.fcnoisy_z_v, on.fcnoisy_z_a
        h = keep["h"]
        gscale = gscale.contiguous().float()
        dzv = torch.empty(B, atoms, device=dev)
        dza = torch.empty(B, A * atoms, device=dev)
        call("riqn_c51_head_bwd", B, A, atoms, ptr(dq), ptr(gscale), ptr(actions), ptr(dzv), ptr(dza))
        dh = torch.empty(B, 2 * hid, device=dev)
        dhv, dha = dh[:, :hid], dh[:, hid:]
        hv, ha = h[:, :hid], h[:, hid:]
        wzv, wza = w_z[:atoms], w_z[atoms:]
        call("riqn_linear_dgrad_ld", B, hid, atoms, ptr(dzv), atoms, ptr(wzv), ptr(dhv), 2 * hid)
        call("riqn_linear_dgrad_ld", B, hid, A * atoms, ptr(dza), A * atoms, ptr(wza), ptr(dha), 2 * hid)
        call("riqn_relu_mask", dh.numel(), ptr(h), ptr(dh))
        scratch = torch.empty(max(A * atoms, 2 * hid), device=dev)
        for layer, d, xin, name in ((zvL, dzv, hv, "fcnoisy_z_v"), (zaL, dza, ha, "fcnoisy_z_a")):
            ew, eb = eps[name]
            call("riqn_noisy_wgrad_ld", B, hid, layer.out_features, ptr(d), layer.out_features, ptr(xin), 2 * hid, ptr(ew),
                 ptr(gv(layer.weight_mu)), ptr(gv(layer.weight_sigma)))
            call("riqn_noisy_bias_grad", B, layer.out_features, ptr(d), ptr(eb), ptr(scratch), ptr(gv(layer.bias_mu)),
                 ptr(gv(layer.bias_sigma)))
        # hidden layers: [h_v | h_a] adjacent in every arena; epsilons of the gradient pass (snapshotted above)
        eps_w_h = torch.cat([eps["fcnoisy_h_v"][0], eps["fcnoisy_h_a"][0]])
        eps_b_h = torch.cat([eps["fcnoisy_h_v"][1], eps["fcnoisy_h_a"][1]])
        call("riqn_noisy_linear_wgrad", B, FEAT, 2 * hid, ptr(dh), ptr(keep["feat"]), ptr(eps_w_h), ptr(eps_b_h), ptr(scratch),
             ptr(gv(hvL.weight_mu)), ptr(gv(hvL.weight_sigma)), ptr(gv(hvL.bias_mu)), ptr(gv(hvL.bias_sigma)))
        dfeat = torch.empty(B, FEAT, device=dev)
        call("riqn_noisy_linear_dgrad", B, FEAT, 2 * hid, ptr(dh), ptr(w_h), ptr(dfeat))
        on.backward_trunk(keep, dfeat)

    return loss, backward


class _C51Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, agent, states, actions, returns, next_states, nonterminals, debug, *params):
        loss, bw = loss_core(agent, states, actions, returns, next_states, nonterminals, debug=debug)
        ctx.bw, ctx.n_params = bw, len(params)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        ctx.bw(grad_loss)
        ctx.bw = None
        return (None,) * (7 + ctx.n_params)


def compute_loss_c51(agent, states, actions, returns, next_states, nonterminals, debug=None):
    if torch.is_grad_enabled():
        params = [p for p in agent.online_net.parameters() if p.requires_grad]
        return _C51Loss.apply(agent, states, actions, returns, next_states, nonterminals, debug, *params)
    loss, _ = loss_core(agent, states, actions, returns, next_states, nonterminals, debug=debug)
    return loss
