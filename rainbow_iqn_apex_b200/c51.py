"""Rainbow-only (C51) branch of the reference: DQN.forward (model.py:120-129) and the categorical loss
(agent.py:77-141), on the same CUDA trunk / NoisyLinear ops as the IQN path plus csrc/c51.cu."""
import torch

from ._lib import call, ptr
from .model import FEAT


def _tc_mode(B):
    """Hidden NoisyLinear products on the tcgen05 path (same arithmetic modes as the IQN head, model.PRECISION)."""
    from .model import PRECISION
    return PRECISION["fwd"] != "fp32" and PRECISION["bwd"] == "bf16" and B % 8 == 0


def forward(net, x, log=False, keep=None, fresh_weights=False, want_argmax=None, support=None):
    """Returns probabilities (or log-probabilities) (B, A, atoms).  model.py:120-129"""
    from .model import PRECISION
    if not fresh_weights:
        net.compose_weights()
    feat = net.trunk(x, keep)
    B = feat.shape[0]
    dev = feat.device
    hid, A, atoms = net.hidden, net.action_space, net.atoms
    h = torch.empty(B, 2 * hid, device=dev)
    x_bf = None
    if _tc_mode(B):
        # features -> 16-bit operand images (fp16 forward: fp16(x) for this product + bf16(x) for the weight gradient)
        fwd = PRECISION["fwd"]
        f16, x3 = fwd == "fp16", fwd == "bf16x3"
        x_hi = torch.empty(B, FEAT, dtype=torch.float16 if f16 else torch.bfloat16, device=dev)
        x_lo = torch.empty(B, FEAT, dtype=torch.bfloat16, device=dev) if (x3 or (f16 and keep is not None)) else None
        call("riqn_split_bf16", B, FEAT, ptr(feat), ptr(x_hi), ptr(x_lo), None, None, 1 if f16 else 0)
        call("riqn_gemm_bf16_tc", B, 2 * hid, FEAT, ptr(x_hi), ptr(x_lo) if x3 else None, ptr(net._w_hi),
             ptr(net._w_lo) if x3 else None, ptr(h), 2 * hid, 1, ptr(net._b_eff_h), None, None, 1, None, None, 3 if f16 else 0)
        x_bf = x_lo if f16 else x_hi
    else:
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
        keep.update(feat=feat, h=h, B=B, x_bf=x_bf)
    return out


def loss_core(agent, states, actions, returns, next_states, nonterminals, debug=None):
    """agent.py:77-141.  Returns (loss (B,), backward(gscale) closure).

    The reference runs online(states) first (:82-83), then the two no-grad passes over next_states (:95-104), each after
    its own reset_noise.  The passes are independent, so they are evaluated here as 2, 3, 1 -- every pass still with its own
    noise sample (injected noises keep their reference slot) -- which leaves the gradient pass's weights and epsilons LIVE
    when the backward runs: no 50 MB of weight / epsilon snapshots per step."""
    from .compute_loss_iqn import _as_device_inputs
    states, actions, returns, next_states, nonterminals = _as_device_inputs(
        agent, states, actions, returns, next_states, nonterminals)
    on, tg = agent.online_net, agent.target_net
    B, A, atoms = states.shape[0], agent.action_space, agent.atoms
    dev = states.device
    inj = getattr(agent, "_inject", None)
    if isinstance(inj, list):
        inj = inj.pop(0) if inj else None
    noises = inj["noises"] if inj else (None, None, None)
    on.reset_noise(noises[1])                                              # :95
    a_star = torch.empty(B, dtype=torch.int64, device=dev)
    forward(on, next_states, fresh_weights=True, want_argmax=a_star, support=agent.support)         # :97-102
    tg.reset_noise(noises[2])                                              # :103
    pns = forward(tg, next_states, fresh_weights=True, support=agent.support)                      # :104
    on.reset_noise(noises[0])                                              # agent.py:82
    keep = {}
    log_ps = forward(on, states, log=True, keep=keep, fresh_weights=True, support=agent.support)   # :83
    loss = torch.empty(B, device=dev)
    dq = torch.empty(B, atoms, device=dev)
    m_out = torch.empty(B, atoms, device=dev) if debug is not None else None
    call("riqn_c51_loss_fwd_bwd", B, A, atoms, ptr(log_ps), ptr(pns), ptr(actions), ptr(a_star), ptr(returns),
         ptr(nonterminals), ptr(agent.support), float(agent.discount ** agent.n), float(agent.Vmin), float(agent.Vmax),
         float(agent.delta_z), ptr(loss), ptr(dq), ptr(m_out))
    if debug is not None:
        debug.update(a_star=a_star, m=m_out, log_ps=log_ps)
    version = getattr(on, "_noise_version", 0)       # bumped by every DQN.reset_noise()

    def backward(gscale, gscale_mul=1.0):
        if getattr(on, "_noise_version", 0) != version:
            raise RuntimeError("the online network's noise was resampled between the C51 loss and its backward")
        hid = on.hidden
        gv = on.grad_view
        hvL, haL, zvL, zaL = on.fcnoisy_h_v, on.fcnoisy_h_a, on.fcnoisy_z_v, on.fcnoisy_z_a
        h = keep["h"]
        gscale = gscale.contiguous().float()
        dzv = torch.empty(B, atoms, device=dev)
        dza = torch.empty(B, A * atoms, device=dev)
        call("riqn_c51_head_bwd", B, A, atoms, ptr(dq), ptr(gscale), float(gscale_mul), ptr(actions), ptr(dzv), ptr(dza))
        dh = torch.empty(B, 2 * hid, device=dev)
        dhv, dha = dh[:, :hid], dh[:, hid:]
        hv, ha = h[:, :hid], h[:, hid:]
        w_z = on._w_eff_z
        wzv, wza = w_z[:atoms], w_z[atoms:]
        call("riqn_linear_dgrad_ld", B, hid, atoms, ptr(dzv), atoms, ptr(wzv), ptr(dhv), 2 * hid)
        call("riqn_linear_dgrad_ld", B, hid, A * atoms, ptr(dza), A * atoms, ptr(wza), ptr(dha), 2 * hid)
        call("riqn_relu_mask", dh.numel(), ptr(h), ptr(dh))
        scratch = torch.empty(max(A * atoms, 2 * hid), device=dev)
        for layer, d, xin in ((zvL, dzv, hv), (zaL, dza, ha)):
            call("riqn_noisy_wgrad_ld", B, hid, layer.out_features, ptr(d), layer.out_features, ptr(xin), 2 * hid,
                 ptr(layer.weight_epsilon), ptr(gv(layer.weight_mu)), ptr(gv(layer.weight_sigma)))
            call("riqn_noisy_bias_grad", B, layer.out_features, ptr(d), ptr(layer.bias_epsilon), ptr(scratch),
                 ptr(gv(layer.bias_mu)), ptr(gv(layer.bias_sigma)))
        # hidden layers: [h_v | h_a] adjacent in every arena (parameters, gradients, epsilons)
        dfeat = torch.empty(B, FEAT, device=dev)
        if keep.get("x_bf") is not None:
            # tensor cores: dW = dh^T x straight from the row-major bf16 images (MN-major operands), dx = dh W from W itself
            from .model import PRECISION
            dh_bf = torch.empty(B, 2 * hid, dtype=torch.bfloat16, device=dev)
            call("riqn_split_bf16", B, 2 * hid, ptr(dh), ptr(dh_bf), None, None, None, 0)
            w_bf = on._w_lo if PRECISION["fwd"] == "fp16" else on._w_hi
            call("riqn_gemm_bf16_tc_mn", 2 * hid, FEAT, B, ptr(dh_bf), ptr(keep["x_bf"]), 1, ptr(gv(hvL.weight_mu)), FEAT, 3,
                 ptr(gv(hvL.weight_sigma)), ptr(hvL.weight_epsilon), 1.0, 1, None, 0)
            call("riqn_noisy_bias_grad", B, 2 * hid, ptr(dh), ptr(hvL.bias_epsilon), ptr(scratch), ptr(gv(hvL.bias_mu)),
                 ptr(gv(hvL.bias_sigma)))
            call("riqn_gemm_bf16_tc_mn", B, FEAT, 2 * hid, ptr(dh_bf), ptr(w_bf), 0, ptr(dfeat), FEAT, 0, None, None, 1.0, 1,
                 None, 0)
        else:
            call("riqn_noisy_linear_wgrad", B, FEAT, 2 * hid, ptr(dh), ptr(keep["feat"]), ptr(hvL.weight_epsilon),
                 ptr(hvL.bias_epsilon), ptr(scratch), ptr(gv(hvL.weight_mu)), ptr(gv(hvL.weight_sigma)), ptr(gv(hvL.bias_mu)),
                 ptr(gv(hvL.bias_sigma)))
            call("riqn_noisy_linear_dgrad", B, FEAT, 2 * hid, ptr(dh), ptr(on._w_eff_h), ptr(dfeat))
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
