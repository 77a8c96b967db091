"""Oracle (test infrastructure): torch-fp32 CPU restatement of the reference network.

Follows ``/root/reference/rainbowiqn/model.py``:
  * NoisyLinear (factorised noise, train/eval forward)      model.py:9-53
  * DQN conv trunk + IQN cosine embedding + dueling head    model.py:56-157
  * C51 (``rainbow_only``) head                             model.py:120-129

Written functionally over a plain ``dict`` of tensors that uses the reference's
``state_dict`` key names, so the same parameter blob can be loaded
into the reference, the oracle and the CUDA product.  Randomness is *injected*:
the caller supplies the already-scaled factor vectors f(eps) and the quantiles tau.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

FEAT = 3136  # 64 * 7 * 7, model.py:118
NOISY_LAYERS = ("fcnoisy_h_v", "fcnoisy_h_a", "fcnoisy_z_v", "fcnoisy_z_a")


def scale_noise(x):
    """f(x) = sign(x) * sqrt(|x|)                              model.py:32-37"""
    return x.sign() * x.abs().sqrt()


def layer_shapes(action_space, hidden=512, embed=64, history=4, rainbow_only=False, atoms=51):
    """Parameter/buffer census of the reference DQN            model.py:59-110, SURVEY §2.2"""
    zv = atoms if rainbow_only else 1
    za = action_space * atoms if rainbow_only else action_space
    shapes = {
        "conv1.weight": (32, history, 8, 8),
        "conv1.bias": (32,),
        "conv2.weight": (64, 32, 4, 4),
        "conv2.bias": (64,),
        "conv3.weight": (64, 64, 3, 3),
        "conv3.bias": (64,),
    }
    if not rainbow_only:
        shapes["iqn_fc.weight"] = (FEAT, embed)
        shapes["iqn_fc.bias"] = (FEAT,)
    for name, (fin, fout) in (
        ("fcnoisy_h_v", (FEAT, hidden)),
        ("fcnoisy_h_a", (FEAT, hidden)),
        ("fcnoisy_z_v", (hidden, zv)),
        ("fcnoisy_z_a", (hidden, za)),
    ):
        shapes[name + ".weight_mu"] = (fout, fin)
        shapes[name + ".weight_sigma"] = (fout, fin)
        shapes[name + ".weight_epsilon"] = (fout, fin)
        shapes[name + ".bias_mu"] = (fout,)
        shapes[name + ".bias_sigma"] = (fout,)
        shapes[name + ".bias_epsilon"] = (fout,)
    return shapes


def is_trainable(key):
    return not key.endswith("_epsilon")


def make_params(seed, action_space=18, hidden=512, embed=64, history=4, noisy_std=0.1,
                rainbow_only=False, atoms=51, sigma_jitter=True):
    """Deterministic parameter blob (numpy RandomState -> identical everywhere).

    Ranges follow the reference initialisers (NoisyLinear.reset_parameters model.py:25-30;
    torch default U(+-1/sqrt(fan_in)) for conv / iqn_fc).  ``sigma_jitter`` perturbs sigma
    so that sigma-gradients are exercised on non-constant values.
    """
    rs = np.random.RandomState(seed)
    out = {}
    for key, shape in layer_shapes(action_space, hidden, embed, history, rainbow_only, atoms).items():
        if key.endswith("_epsilon"):
            arr = np.zeros(shape, np.float32)
        elif key.endswith("weight_sigma"):
            base = noisy_std / math.sqrt(shape[1])
            arr = np.full(shape, base, np.float32)
            if sigma_jitter:
                arr = (arr * rs.uniform(0.5, 1.5, shape)).astype(np.float32)
        elif key.endswith("bias_sigma"):
            base = noisy_std / math.sqrt(shape[0])
            arr = np.full(shape, base, np.float32)
            if sigma_jitter:
                arr = (arr * rs.uniform(0.5, 1.5, shape)).astype(np.float32)
        else:
            if key.endswith("bias_mu"):
                fan_in = shapes_fan_in(key, action_space, hidden, rainbow_only, atoms)
            elif key.endswith(".bias"):
                fan_in = {"conv1.bias": history * 64, "conv2.bias": 32 * 16, "conv3.bias": 64 * 9,
                          "iqn_fc.bias": embed}[key]
            else:
                fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            arr = rs.uniform(-bound, bound, shape).astype(np.float32)
        out[key] = arr
    return out


def shapes_fan_in(key, action_space, hidden, rainbow_only, atoms):
    return FEAT if "fcnoisy_h_" in key else hidden


def to_torch(params, requires_grad=False):
    out = {}
    for k, v in params.items():
        t = torch.from_numpy(np.ascontiguousarray(v)).clone() if isinstance(v, np.ndarray) else v.clone()
        if requires_grad and is_trainable(k):
            t.requires_grad_(True)
        out[k] = t
    return out


def make_noise(seed, action_space=18, hidden=512, rainbow_only=False, atoms=51):
    """One reset's worth of scaled factor vectors, in the reference's draw order:
    for each noisy child in ``named_children`` order, f(eps_in) then f(eps_out)
    (DQN.reset_noise model.py:159-162, NoisyLinear.reset_noise model.py:39-43)."""
    rs = np.random.RandomState(seed)
    zv = atoms if rainbow_only else 1
    za = action_space * atoms if rainbow_only else action_space
    dims = {"fcnoisy_h_v": (FEAT, hidden), "fcnoisy_h_a": (FEAT, hidden),
            "fcnoisy_z_v": (hidden, zv), "fcnoisy_z_a": (hidden, za)}
    noise = {}
    for name in NOISY_LAYERS:
        fin, fout = dims[name]
        e_in = scale_noise(torch.from_numpy(rs.standard_normal(fin).astype(np.float32)))
        e_out = scale_noise(torch.from_numpy(rs.standard_normal(fout).astype(np.float32)))
        noise[name] = (e_in, e_out)
    return noise


def apply_noise(p, noise):
    """NoisyLinear.reset_noise with injected factors: eps_w = eps_out (x) eps_in, eps_b = eps_out
    (model.py:39-43).  Mutates and returns ``p``."""
    for name, (e_in, e_out) in noise.items():
        p[name + ".weight_epsilon"] = torch.outer(e_out, e_in)
        p[name + ".bias_epsilon"] = e_out.clone()
    return p


def noisy_linear(p, name, x, training=True):
    """NoisyLinear.forward                                     model.py:45-53"""
    if training:
        w = p[name + ".weight_mu"] + p[name + ".weight_sigma"] * p[name + ".weight_epsilon"]
        b = p[name + ".bias_mu"] + p[name + ".bias_sigma"] * p[name + ".bias_epsilon"]
        return F.linear(x, w, b)
    return F.linear(x, p[name + ".weight_mu"], p[name + ".bias_mu"])


def conv_trunk(p, x, keep=None):
    """conv1(8x8,s4,p1) conv2(4x4,s2) conv3(3x3) + ReLU, flatten C-major   model.py:65-67,115-118"""
    o1 = F.relu(F.conv2d(x, p["conv1.weight"], p["conv1.bias"], stride=4, padding=1))
    o2 = F.relu(F.conv2d(o1, p["conv2.weight"], p["conv2.bias"], stride=2))
    o3 = F.relu(F.conv2d(o2, p["conv3.weight"], p["conv3.bias"]))
    if keep is not None:
        keep.update(o1=o1, o2=o2, o3=o3)
    return o3.reshape(-1, FEAT)


def cos_embedding(tau, embed):
    """cos(fl(fl(i)*fl(pi)) * tau), i = 1..embed                model.py:136-144"""
    tiled = tau.repeat([1, embed])
    return torch.cos(torch.arange(1, embed + 1, 1, dtype=torch.float32) * math.pi * tiled)


def dqn_forward_iqn(p, x, num_quantiles, tau, training=True, keep=None):
    """DQN.forward, IQN branch                                  model.py:112-118,130-157

    ``x`` (B,4,84,84) fp32 in [0,1]; ``tau`` (num_quantiles*B, 1) fp32, row = q*B + b.
    Returns q (num_quantiles*B, A).  ``keep`` (a dict) receives intermediates.
    """
    batch = x.shape[0]
    feat = conv_trunk(p, x, keep)
    embed = p["iqn_fc.weight"].shape[1]
    cosv = cos_embedding(tau, embed)
    phi = F.relu(F.linear(cosv, p["iqn_fc.weight"], p["iqn_fc.bias"]))
    xt = feat.repeat(num_quantiles, 1) * phi
    h_v = F.relu(noisy_linear(p, "fcnoisy_h_v", xt, training))
    h_a = F.relu(noisy_linear(p, "fcnoisy_h_a", xt, training))
    v = noisy_linear(p, "fcnoisy_z_v", h_v, training)
    a = noisy_linear(p, "fcnoisy_z_a", h_a, training)
    q = v + a - a.mean(1, keepdim=True)
    if keep is not None:
        keep.update(feat=feat, cos=cosv, phi=phi, x=xt, h_v=h_v, h_a=h_a, v=v, a=a, q=q)
    assert q.shape[0] == num_quantiles * batch
    return q


def dqn_forward_c51(p, x, action_space, atoms, log=False, training=True):
    """DQN.forward, rainbow_only branch                         model.py:120-129"""
    feat = conv_trunk(p, x)
    v = noisy_linear(p, "fcnoisy_z_v", F.relu(noisy_linear(p, "fcnoisy_h_v", feat, training)), training)
    a = noisy_linear(p, "fcnoisy_z_a", F.relu(noisy_linear(p, "fcnoisy_h_a", feat, training)), training)
    v, a = v.view(-1, 1, atoms), a.view(-1, action_space, atoms)
    q = v + a - a.mean(1, keepdim=True)
    return F.log_softmax(q, dim=2) if log else F.softmax(q, dim=2)
