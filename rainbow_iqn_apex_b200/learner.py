"""Learner -- mirror of the reference ``rainbowiqn/learner.py:8-36``.

``learn(mem, mp_queue) -> (idxs, loss)`` performs the reference's sequence
sample -> loss -> zero_grad -> (weights*loss).mean().backward() -> Adam.step (learner.py:14-26) with the
backward driven directly (no autograd graph) and the IS weights folded into the upstream gradient
gscale[b] = weights[b] / B.  ``north_star`` spellings Agent.learn / Agent.update_target are aliased.
"""
import io

import torch

from . import compute_loss_iqn
from .agent import Agent

MODEL_WEIGHT_STR = "model_weight"      # rainbowiqn/constants.py:18
STEP_LEARNER_STR = "step_learner:"     # rainbowiqn/constants.py:14


class Learner(Agent):
    def __init__(self, args, action_space, redis_servor):
        super().__init__(args, action_space, redis_servor)
        self.process_group = None  # set by parallel.DataParallelLearner

    def learn(self, mem_redis, mp_queue):
        sample = mem_redis.get_sample_from_mp_queue(mp_queue)
        idxs, states, actions, returns, next_states, nonterminals, weights = sample
        loss = self.learn_on_batch(states, actions, returns, next_states, nonterminals, weights)
        return idxs, loss

    def learn_on_batch(self, states, actions, returns, next_states, nonterminals, weights):
        """learner.py:18-24 on an already assembled minibatch.  Returns the per-transition loss (B,)."""
        on = self.online_net
        dev = on._flat.device
        weights = weights.to(dev, torch.float32)
        if self.rainbow_only:
            from . import c51
            loss, bw = c51.loss_core(self, states, actions, returns, next_states, nonterminals)
            on.zero_grad()
            bw(weights / weights.shape[0])
        else:
            loss, dtheta, keep, actions = compute_loss_iqn.loss_core(
                self, states, actions, returns, next_states, nonterminals, keep_graph=True)
            on.zero_grad()                                                      # learner.py:22
            on.backward_iqn(keep, dtheta, weights / weights.shape[0], actions)  # learner.py:23
        if self.process_group is not None:
            torch.distributed.all_reduce(on._flat_grad, group=self.process_group)
        self.optimiser.step()                                                   # learner.py:24
        return loss

    # north_star spellings
    update_target = Agent.update_target_net

    def save_to_redis(self, T_learner):
        """learner.py:28-36 (kept for wire compatibility; needs a redis-like object with pipeline())."""
        save_bytesIO = io.BytesIO()
        torch.save(self.online_net.state_dict(), save_bytesIO)
        pipe = self.redis_servor.pipeline()
        pipe.set(MODEL_WEIGHT_STR, save_bytesIO.getvalue())
        pipe.set(STEP_LEARNER_STR, T_learner)
        pipe.execute()
