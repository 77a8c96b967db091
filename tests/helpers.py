"""Shared helpers for the parity tests (tests may import oracle/; the product never does)."""
from types import SimpleNamespace

import numpy as np
import torch


def make_args(device, batch=32, cfg=None, rainbow_only=False, nb_actor=1, actor_capacity=1000):
    cfg = cfg or {}
    return SimpleNamespace(
        multi_step=cfg.get("n_step", 3), history_length=4, discount=cfg.get("discount", 0.99), device=device,
        batch_size=batch, length_actor_buffer=1000, model=None, lr=6.25e-5 if rainbow_only else 5e-5,
        adam_eps=1.5e-4 if rainbow_only else 3.125e-4, rainbow_only=int(rainbow_only), atoms=51, V_min=-10.0,
        V_max=10.0, kappa=cfg.get("kappa", 1.0), num_tau_samples=cfg.get("n_tau", 64),
        num_tau_prime_samples=cfg.get("n_tau_prime", 64), num_quantile_samples=cfg.get("n_quantile", 32),
        quantile_embedding_dim=64, hidden_size=512, noisy_std=0.1, disable_cuda=False, nb_actor=nb_actor,
        actor_capacity=actor_capacity, priority_weight=0.4, priority_exponent=0.2)


def load_params(net, params):
    """Load a numpy parameter blob (oracle.network.make_params) into a DQN through load_state_dict."""
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()}
    net.load_state_dict(sd)
    net.compose_weights()


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def digest(t, head=8):
    a = t.detach().cpu().numpy().astype(np.float64).ravel()
    return np.concatenate([[a.sum(), np.abs(a).sum(), np.sqrt((a * a).sum())], a[:head]])
