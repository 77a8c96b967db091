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
    def __init__(self, args, action_space, redis_servor):
        _lib.require_device()
        self.action_space = action_space
        self.n = args.multi_step
        self.history = args.history_length
        self.discount = args.discount
        self.redis_servor = redis_servor
        self.device = args.device
        self.batch_size = args.batch_size
        self.length_actor_buffer = getattr(args, "length_actor_buffer", 1000)

        self.online_net = DQN(args, self.action_space).to(device=args.device)
        checkpoint = None
        if getattr(args, "model", None):
            if os.path.isfile(args.model):
                print("We loaded model ", args.model)
                checkpoint = torch.load(args.model, map_location="cpu")
                self.online_net.load_state_dict(checkpoint["model_state_dict"])
            else:
                print("We didn't fint the model you gave as input!")
                raise Exception
        self.online_net.train()

        self.target_net = DQN(args, self.action_space).to(device=args.device)
        self.update_target_net()
        self.target_net.train()                      # target stays noisy (agent.py:37-41)
        for param in self.target_net.parameters():
            param.requires_grad = False

        self.optimiser = Adam(self.online_net.parameters(), lr=args.lr, eps=args.adam_eps)
        if checkpoint is not None:
            self.optimiser.load_state_dict(checkpoint["optimiser_state_dict"])

        self.rainbow_only = args.rainbow_only
        if self.rainbow_only:
            self.atoms = args.atoms
            self.Vmin = args.V_min
            self.Vmax = args.V_max
            self.support = torch.linspace(args.V_min, args.V_max, self.atoms).to(device=args.device)
            self.delta_z = (args.V_max - args.V_min) / (self.atoms - 1)
        else:
            self.kappa = args.kappa
            self.num_tau_samples = args.num_tau_samples
            self.num_tau_prime_samples = args.num_tau_prime_samples
            self.num_quantile_samples = args.num_quantile_samples
        self._inject = None  # parity hook: {"noises": (n0, n1, n2), "taus": (t0, t1, t2)}

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
