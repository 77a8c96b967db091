"""Oracle (test infrastructure): restatement of the reference actor-side computations.

Follows
  * ``Actor.act``                  /root/reference/rainbowiqn/actor.py:15-25
  * ``Actor.compute_priorities``   rainbowiqn/actor.py:41-124
  * the buffer flush of the actor loop (initial priorities + the ``max_priority`` tail for the last n steps)
                                   rainbowiqn/launch_actor.py:116-133

Randomness is injected like in ``oracle/losses.py``: one noise dict per ``reset_noise`` and one tau array per
``DQN.forward`` call, in call order.
"""
import math

import numpy as np
import torch

from . import losses, network as net


def act(p_online, state_buffer, n_quantile, tau):
    """actor.py:15-25 (IQN branch): greedy action of the mean over K sampled quantiles.  ``state_buffer`` is the
    list of ``history`` (84, 84) uint8 frames; the caller has applied the noise it wants (launch_actor.py:76-77).
    Returns (action, q_mean (A,))."""
    state = torch.from_numpy(np.stack(state_buffer).astype(np.float32) / 255)
    with torch.no_grad():
        q = net.dqn_forward_iqn(p_online, state.unsqueeze(0), n_quantile, tau)
        q_mean = q.mean(0)
    return int(q_mean.argmax(0)), q_mean


def compute_priorities(p_online, p_target, tab_state, tab_action, tab_reward, tab_nonterminal, priority_exponent,
                       noises, taus, cfg, batch_size, history=4):
    """actor.py:41-124.  ``noises`` / ``taus``: one (3 noise dicts) / (3 tau arrays) tuple per chunk of
    ``batch_size`` transitions, in order.  Returns loss ** priority_exponent (len(tab_action) - n,)."""
    n, discount = cfg["n_step"], cfg["discount"]
    len_buffer = len(tab_action)
    assert len(tab_action) == len(tab_reward) == len(tab_nonterminal) == len(tab_state) - history + 1
    nonterm = np.float32(tab_nonterminal[n:])
    for indice in np.where(nonterm == 0)[0]:                                   # :67-69
        nonterm[indice + 1:(indice + n + 1)] = 0
    actions = torch.tensor(tab_action[:len_buffer - n], dtype=torch.int64)
    tab_returns = [sum(discount ** k * tab_reward[k + i] for k in range(n)) for i in range(0, len_buffer - n)]
    returns = torch.tensor(tab_returns, dtype=torch.float32)
    nonterminals = torch.tensor(nonterm, dtype=torch.float32)
    out = []
    for c in range(math.ceil(len(actions) / batch_size)):
        lo, hi = c * batch_size, min((c + 1) * batch_size, len(actions))
        states = torch.from_numpy(np.stack([np.stack(tab_state[i:i + history]) for i in range(lo, hi)])).float().div_(255)
        nexts = torch.from_numpy(np.stack([np.stack(tab_state[i + n:i + history + n]) for i in range(lo, hi)])).float().div_(255)
        with torch.no_grad():
            loss = losses.iqn_loss(p_online, p_target, states, actions[lo:hi], returns[lo:hi], nexts, nonterminals[lo:hi],
                                   noises[c], taus[c], **cfg)
        out.append(loss.numpy())
    return np.power(np.concatenate(out), priority_exponent)


def flush_priorities(priorities_buffer, max_priority, n):
    """launch_actor.py:127-133: the last n steps of a flushed buffer have no next_state yet; they enter the replay
    with the current max priority."""
    return np.concatenate((priorities_buffer, np.ones(n) * np.float64(max_priority)))
