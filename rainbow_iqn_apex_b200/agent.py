"""Agent -- mirror of the reference ``rainbowiqn/agent.py:10-166`` (shared by Learner and Actor).

Same constructor ``Agent(args, action_space, redis_servor)`` reading the same ``args`` fields, same public
attributes (online_net, target_net, optimiser, n, history, discount, device, batch_size, kappa, num_tau_samples,
num_tau_prime_samples, num_quantile_samples, support, ...) and methods (reset_noise, update_target_net,
compute_loss_actor_or_learner, save, train, eval).  The networks are rainbow_iqn_apex_b200.model.DQN (CUDA) and the
optimiser is the arena Adam; checkpoints keep the reference schema
{T_actors, T_learner, model_state_dict, optimiser_state_dict} (agent.py:150-160).
"""
import os

import torch

from . import _lib
from . import compute_loss_iqn
from .model import DQN
from .optim import Adam


class Agent:
    # public attribute <- args field (agent.py:13-63; read by the loss code, the actors and the launch scripts)
    _ARG_FIELDS = (("n", "multi_step"), ("history", "history_length"), ("discount", "discount"), ("device", "device"),
                   ("batch_size", "batch_size"), ("rainbow_only", "rainbow_only"))
    _C51_FIELDS = (("atoms", "atoms"), ("Vmin", "V_min"), ("Vmax", "V_max"))
    _IQN_FIELDS = ("kappa", "num_tau_samples", "num_tau_prime_samples", "num_quantile_samples")

    def __init__(self, args, action_space, redis_servor):
        _lib.require_device()
        self.action_space, self.redis_servor = action_space, redis_servor
        for attr, field in self._ARG_FIELDS:
            setattr(self, attr, getattr(args, field))
        self.length_actor_buffer = getattr(args, "length_actor_buffer", 1000)

        # online network (+ optional checkpoint, agent.py:26-34), noisy target copy with frozen parameters (:37-41), Adam (:43)
        checkpoint = self._read_checkpoint(getattr(args, "model", None))
        self.online_net = DQN(args, action_space).to(device=args.device)
        if checkpoint is not None:
            self.online_net.load_state_dict(checkpoint["model_state_dict"])
        self.target_net = DQN(args, action_space).to(device=args.device)
        self.update_target_net()
        for net in (self.online_net, self.target_net):
            net.train()                                  # the target stays in train mode: it is noisy too
        for p in self.target_net.parameters():
            p.requires_grad = False
        self.optimiser = Adam(self.online_net.parameters(), lr=args.lr, eps=args.adam_eps)
        if checkpoint is not None:
            self.optimiser.load_state_dict(checkpoint["optimiser_state_dict"])

        if self.rainbow_only:                            # categorical support (agent.py:49-57)
            for attr, field in self._C51_FIELDS:
                setattr(self, attr, getattr(args, field))
            self.support = torch.linspace(self.Vmin, self.Vmax, self.atoms).to(device=args.device)
            self.delta_z = (self.Vmax - self.Vmin) / (self.atoms - 1)
        else:                                            # IQN sampling sizes (agent.py:58-63)
            for field in self._IQN_FIELDS:
                setattr(self, field, getattr(args, field))
        self._inject = None  # parity hook: {"noises": (n0, n1, n2), "taus": (t0, t1, t2)}

    @staticmethod
    def _read_checkpoint(path):
        """agent.py:26-34: a given but missing checkpoint is an error (bare Exception, like the reference)."""
        if not path:
            return None
        if not os.path.isfile(path):
            print("We didn't fint the model you gave as input!")
            raise Exception
        print("We loaded model ", path)
        return torch.load(path, map_location="cpu")

    def reset_noise(self):
        """agent.py:66-67"""
        self.online_net.reset_noise()

    def update_target_net(self):
        """agent.py:69-70 -- parameters AND epsilon buffers, as load_state_dict(state_dict()) copies them;
        here two flat device copies instead of 32 tensor copies."""
        self.target_net._flat.copy_(self.online_net._flat)
        self.target_net._eps_flat.copy_(self.online_net._eps_flat)
        self.target_net.compose_weights()

    def compute_loss_actor_or_learner(self, states, actions, returns, next_states, nonterminals, debug=None):
        """agent.py:72-147"""
        if self.rainbow_only:
            from . import c51
            return c51.compute_loss_c51(self, states, actions, returns, next_states, nonterminals, debug=debug)
        return compute_loss_iqn.compute_loss_actor_or_learner_iqn(
            self, states, actions, returns, next_states, nonterminals, debug=debug)

    def save(self, path, T_actors, T_learner, name):
        """agent.py:150-160"""
        torch.save(
            {
                "T_actors": T_actors,
                "T_learner": T_learner,
                "model_state_dict": self.online_net.state_dict(),
                "optimiser_state_dict": self.optimiser.state_dict(),
            },
            os.path.join(path, name),
        )

    def train(self):
        self.online_net.train()

    def eval(self):
        self.online_net.eval()
