"""Oracle (test infrastructure): seeded synthetic inputs shared by the golden generator, the
parity tests and the benchmark (SURVEY.md §8d "Synthetic inputs")."""
import numpy as np
import torch

from . import network as net


def iqn_cfg(n_tau=64, n_tau_prime=64, n_quantile=32, discount=0.99, n_step=3, kappa=1.0):
    return dict(n_tau=n_tau, n_tau_prime=n_tau_prime, n_quantile=n_quantile, discount=discount,
                n_step=n_step, kappa=kappa)


def make_batch(seed, batch, action_space=18, n_step=3, discount=0.99):
    """uint8 frame stacks + actions/returns/nonterminals/IS-weights (numpy)."""
    rs = np.random.RandomState(seed)
    states = rs.randint(0, 256, (batch, 4, 84, 84)).astype(np.uint8)
    next_states = rs.randint(0, 256, (batch, 4, 84, 84)).astype(np.uint8)
    actions = rs.randint(0, action_space, batch).astype(np.int64)
    rewards = rs.randint(-1, 2, (batch, n_step)).astype(np.float64)
    returns = sum(discount ** k * rewards[:, k] for k in range(n_step)).astype(np.float32)
    nonterminals = (rs.uniform(size=batch) < 0.9).astype(np.float32)
    weights = rs.uniform(0.1, 1.0, batch).astype(np.float32)
    return dict(states=states, next_states=next_states, actions=actions, returns=returns,
                nonterminals=nonterminals, weights=weights)


def make_taus(seed, batch, cfg):
    rs = np.random.RandomState(seed)
    return tuple(rs.uniform(0, 1, (nq * batch, 1)).astype(np.float32)
                 for nq in (cfg["n_quantile"], cfg["n_tau_prime"], cfg["n_tau"]))


def make_noises(seed, count=3, **kw):
    return tuple(net.make_noise(seed + 1000 * i, **kw) for i in range(count))


def batch_to_torch(b):
    """What the reference hands the loss: fp32 frames / 255        redis_memory.py:527-536"""
    return (torch.from_numpy(b["states"]).to(torch.float32).div_(255),
            torch.from_numpy(b["actions"]),
            torch.from_numpy(b["returns"]),
            torch.from_numpy(b["next_states"]).to(torch.float32).div_(255),
            torch.from_numpy(b["nonterminals"]))


def tensor_digest(t, head=8):
    """Small fingerprint of a tensor for golden files: [sum, abs-sum, l2] (float64) + first elements."""
    a = t.detach().cpu().numpy().astype(np.float64).ravel() if hasattr(t, "detach") else np.asarray(t, np.float64).ravel()
    return np.concatenate([[a.sum(), np.abs(a).sum(), np.sqrt((a * a).sum())], a[:head]])
