"""Golden-vector generator + oracle pin (dev container only; needs /root/reference).

    python -m oracle.make_golden            # writes tests/golden/*.npz

Imports the UNMODIFIED reference modules from /root/reference (never copied into this repo),
runs them with injected noise / quantiles / sample values, asserts that the oracle restatement
(``oracle/*.py``) reproduces them, and stores small fixtures that travel to the GPU box.

How the reference is made runnable offline:
  * ``redlock`` is absent   -> a stub module with Redlock.lock/unlock is placed in sys.modules;
  * no redis server         -> an in-process fake StrictRedis that stores ``repr(float)`` strings;
  * ``ndarray.tostring`` is gone in numpy 2 -> frames are wrapped in an object offering
    ``.ravel().tostring()``;
  * NoisyLinear._scale_noise (model.py:32) and torch.FloatTensor(...).uniform_ (model.py:132) are
    patched to pop injected vectors, so oracle / reference / CUDA runs see identical randomness.
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

from . import actor as oactor, cases, losses, network as net, replay as oreplay, sumtree as osum  # noqa: E402


# --------------------------------------------------------------------------- reference harness
def _install_stubs():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "redlock" not in sys.modules:
        m = types.ModuleType("redlock")

        class Redlock:
            def __init__(self, *a, **k):
                pass

            def lock(self, *a, **k):
                return True

            def unlock(self, *a, **k):
                return True

        m.Redlock = Redlock
        sys.modules["redlock"] = m


class FakePipe:
    def __init__(self, r):
        self.r, self.ops = r, []

    def __getattr__(self, name):
        def rec(*a, **k):
            self.ops.append((name, a, k))
            return self
        return rec

    def execute(self):
        out = [getattr(self.r, name)(*a, **k) for name, a, k in self.ops]
        self.ops = []
        return out


class FakeRedis:
    """Just enough of redis.StrictRedis; numbers stored as decimal strings like the server."""

    def __init__(self):
        self.kv = {}

    @staticmethod
    def _enc(v):
        if isinstance(v, bytes):
            return v
        if isinstance(v, (float, np.floating)):
            return repr(float(v)).encode()
        if isinstance(v, (int, np.integer)):
            return str(int(v)).encode()
        return str(v).encode()

    def flushdb(self):
        self.kv.clear()

    def set(self, k, v):
        self.kv[k] = self._enc(v)

    def get(self, k):
        return self.kv.get(k)

    def incrbyfloat(self, k, v):
        cur = float(self.kv.get(k, b"0"))
        self.kv[k] = repr(cur + float(v)).encode()

    def hmset(self, k, d):
        self.kv[k] = {f: self._enc(v) for f, v in d.items()}

    def hmget(self, k, *fields):
        h = self.kv.get(k)
        return [None] * len(fields) if h is None else [h.get(f) for f in fields]

    def pipeline(self):
        return FakePipe(self)


class _Frame:
    """numpy-2 shim for ``np_state.ravel().tostring()`` (redis_memory.py:182)."""

    def __init__(self, a):
        self.a = a

    def ravel(self):
        return self

    def tostring(self):
        return self.a.tobytes()


def ref_args(batch, cfg, rainbow_only=False, nb_actor=1, actor_capacity=1000):
    return SimpleNamespace(
        multi_step=cfg["n_step"], history_length=4, discount=cfg["discount"], device=torch.device("cpu"),
        batch_size=batch, length_actor_buffer=1000, model=None, lr=5e-5 if not rainbow_only else 6.25e-5,
        adam_eps=3.125e-4 if not rainbow_only else 1.5e-4, rainbow_only=int(rainbow_only), atoms=51,
        V_min=-10.0, V_max=10.0, kappa=cfg.get("kappa", 1.0), num_tau_samples=cfg.get("n_tau", 64),
        num_tau_prime_samples=cfg.get("n_tau_prime", 64), num_quantile_samples=cfg.get("n_quantile", 32),
        quantile_embedding_dim=64, hidden_size=512, noisy_std=0.1, disable_cuda=True,
        actor_capacity=actor_capacity, nb_actor=nb_actor, priority_weight=0.4, priority_exponent=0.2,
        host_redis="localhost", port_redis=6379, synchronize_actors_with_learner=1)


class Injector:
    """Feeds injected noise factors and quantiles to the unmodified reference model."""

    def __init__(self):
        self.noise_q, self.tau_q = [], []

    def push_noise(self, noise):
        for name in net.NOISY_LAYERS:
            e_in, e_out = noise[name]
            self.noise_q += [e_in.clone(), e_out.clone()]

    def push_tau(self, tau):
        self.tau_q.append(torch.from_numpy(np.ascontiguousarray(tau)).clone())

    def __enter__(self):
        import rainbowiqn.model as rmodel
        self._old_scale = rmodel.NoisyLinear._scale_noise
        self._old_ft = torch.FloatTensor
        inj = self

        def scale(self_layer, size):
            v = inj.noise_q.pop(0)
            assert v.numel() == size, (v.numel(), size)
            return v

        class FakeFloatTensor:
            def __new__(cls, *size):
                t = inj.tau_q.pop(0)
                assert tuple(t.shape) == tuple(size), (t.shape, size)
                return SimpleNamespace(uniform_=lambda a, b: t)

        rmodel.NoisyLinear._scale_noise = scale
        torch.FloatTensor = FakeFloatTensor
        return self

    def __exit__(self, *exc):
        import rainbowiqn.model as rmodel
        rmodel.NoisyLinear._scale_noise = self._old_scale
        torch.FloatTensor = self._old_ft


def build_ref_learner(params, batch, cfg, rainbow_only=False):
    from rainbowiqn.learner import Learner
    inj = Injector()
    # constructor resets noise of both nets once (model.py:23); feed throwaway factors
    for _ in range(2):
        inj.push_noise(net.make_noise(99, rainbow_only=rainbow_only))
    with inj:
        learner = Learner(ref_args(batch, cfg, rainbow_only), 18, None)
    sd = {k: torch.from_numpy(v.copy()) for k, v in params.items()}
    learner.online_net.load_state_dict(sd)
    learner.update_target_net()
    learner.train()
    return learner


class _FakeMem:
    def __init__(self, sample):
        self.sample = sample

    def get_sample_from_mp_queue(self, q):
        return self.sample


# --------------------------------------------------------------------------- golden cases
def golden_iqn(name, batch, cfg, steps, seed):
    """Run ``steps`` reference Learner.learn calls and the oracle side by side."""
    params = net.make_params(seed)
    learner = build_ref_learner(params, batch, cfg)
    p_on = net.to_torch(params, requires_grad=True)
    p_tg = net.to_torch(params)
    adam = losses.Adam([k for k in p_on if net.is_trainable(k)], lr=5e-5, eps=3.125e-4)
    rec = {"batch": batch, "steps": steps, "seed": seed, **{f"cfg_{k}": v for k, v in cfg.items()}}
    for s in range(steps):
        b = cases.make_batch(seed + 10 + s, batch, n_step=cfg["n_step"], discount=cfg["discount"])
        taus = cases.make_taus(seed + 20 + s, batch, cfg)
        noises = cases.make_noises(seed + 30 + s)
        tb = cases.batch_to_torch(b)
        w = torch.from_numpy(b["weights"])
        inj = Injector()
        for nz in noises:
            inj.push_noise(nz)
        for t in taus:
            inj.push_tau(t)
        with inj:
            idxs, ref_loss = learner.learn(_FakeMem((np.arange(batch), tb[0], tb[1], tb[2], tb[3], tb[4], w)), None)
        assert not inj.noise_q and not inj.tau_q
        keep = {}
        o_loss, o_grads = losses.learn_step(p_on, p_tg, adam, tb, w, noises,
                                            tuple(torch.from_numpy(t) for t in taus), cfg, keep=keep)
        ref_loss = ref_loss.detach()
        err = float((ref_loss - o_loss).abs().max() / ref_loss.abs().max())
        print(f"[{name}] step {s}: loss max-rel-diff oracle vs reference = {err:.3e}")
        assert err < 1e-5, err
        for k, g in o_grads.items():
            rg = dict(learner.online_net.named_parameters())[k].grad
            gerr = float((rg - g).abs().max() / (rg.abs().max() + 1e-30))
            assert gerr < 1e-4, (k, gerr)
        for k, t in learner.online_net.state_dict().items():
            if net.is_trainable(k):
                perr = float((t - p_on[k].detach()).abs().max())
                assert perr < 1e-6, (k, perr)
        rec[f"loss_{s}"] = ref_loss.numpy()
        rec[f"a_star_{s}"] = keep["a_star"].numpy()
        rec[f"target_{s}"] = keep["target"].numpy()
        rec[f"theta_{s}"] = keep["theta"].detach().numpy()
        for k in o_grads:
            rg = dict(learner.online_net.named_parameters())[k].grad
            rec[f"grad_{s}_{k}"] = cases.tensor_digest(rg)
            rec[f"param_{s}_{k}"] = cases.tensor_digest(learner.online_net.state_dict()[k])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)


def golden_c51(name, batch, steps, seed):
    cfg = dict(discount=0.99, n_step=3)
    ocfg = dict(atoms=51, v_min=-10.0, v_max=10.0, discount=0.99, n_step=3)
    params = net.make_params(seed, rainbow_only=True)
    learner = build_ref_learner(params, batch, cfg, rainbow_only=True)
    p_on = net.to_torch(params, requires_grad=True)
    p_tg = net.to_torch(params)
    adam = losses.Adam([k for k in p_on if net.is_trainable(k)], lr=6.25e-5, eps=1.5e-4)
    rec = {"batch": batch, "steps": steps, "seed": seed}
    for s in range(steps):
        b = cases.make_batch(seed + 10 + s, batch)
        noises = cases.make_noises(seed + 30 + s, rainbow_only=True)
        tb = cases.batch_to_torch(b)
        w = torch.from_numpy(b["weights"])
        inj = Injector()
        for nz in noises:
            inj.push_noise(nz)
        with inj:
            _, ref_loss = learner.learn(_FakeMem((np.arange(batch), tb[0], tb[1], tb[2], tb[3], tb[4], w)), None)
        keep = {}
        o_loss, o_grads = losses.learn_step(p_on, p_tg, adam, tb, w, noises, None, ocfg, rainbow_only=True, keep=keep)
        ref_loss = ref_loss.detach()
        err = float((ref_loss - o_loss).abs().max() / ref_loss.abs().max())
        print(f"[{name}] step {s}: loss max-rel-diff oracle vs reference = {err:.3e}")
        assert err < 1e-5, err
        rec[f"loss_{s}"] = ref_loss.numpy()
        rec[f"a_star_{s}"] = keep["a_star"].numpy()
        rec[f"m_{s}"] = keep["m"].numpy()
        for k in o_grads:
            rg = dict(learner.online_net.named_parameters())[k].grad
            rec[f"grad_{s}_{k}"] = cases.tensor_digest(rg)
            rec[f"param_{s}_{k}"] = cases.tensor_digest(learner.online_net.state_dict()[k])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)


def _ref_tree_array(fr, cap):
    return np.array([float(fr.get("priorities:" + str(i))) for i in range(2 * cap - 1)], np.float64)


def golden_sumtree(name, actor_capacity, nb_actor, batch, rounds, seed):
    """append -> (find, weights, update)* through the unmodified RedisSegmentTree / ReplayRedisMemory."""
    import rainbowiqn.redis_memory as rm
    rs = np.random.RandomState(seed)
    cap = actor_capacity * nb_actor
    fr = FakeRedis()
    args = ref_args(batch, cases.iqn_cfg(), nb_actor=nb_actor, actor_capacity=actor_capacity)
    mem = rm.ReplayRedisMemory(args, fr)
    mem.transitions.initialize_redis_database()
    tree = osum.SumTree(actor_capacity, nb_actor)
    store = oreplay.ReplayStore(actor_capacity, nb_actor)
    rec = {"actor_capacity": actor_capacity, "nb_actor": nb_actor, "batch": batch, "rounds": rounds, "seed": seed}
    # fill every actor segment completely, in 2 chunks each, write heads end at random positions
    heads = []
    for a in range(nb_actor):
        start = 0
        chunks = [actor_capacity // 2, actor_capacity - actor_capacity // 2, int(rs.randint(5, actor_capacity // 3))]
        for ci, n in enumerate(chunks):
            ts = (np.arange(n) + (0 if ci == 0 else 7)).astype(np.int64)
            if ci == 0:
                ts[0] = 0
            ts[n // 2] = 0  # an episode start in the middle
            frame_seed = seed * 1000 + a * 10 + ci
            frames = np.random.RandomState(frame_seed).randint(0, 256, (n, 84, 84)).astype(np.uint8)
            actions = rs.randint(0, 18, n)
            rewards = rs.randint(-1, 2, n).astype(np.float64)
            dones = rs.uniform(size=n) < 0.05
            dones[n // 2 - 1] = True
            pri = (rs.uniform(0.01, 1.0, n) ** 0.2).astype(np.float32)
            buf = [[int(ts[i]), _Frame(frames[i]), int(actions[i]), float(rewards[i]), bool(dones[i])] for i in range(n)]
            mem.transitions.append_actor_buffer(buf, start, a, pri, 0)
            tree.append_priorities(start, a, pri)
            store.write(a, start, ts, frames, actions, rewards, dones)
            rec[f"append_{a}_{ci}"] = np.array([start, n])
            rec[f"append_pri_{a}_{ci}"] = pri
            rec[f"append_ts_{a}_{ci}"] = ts
            rec[f"append_act_{a}_{ci}"] = actions
            rec[f"append_rew_{a}_{ci}"] = rewards
            rec[f"append_done_{a}_{ci}"] = dones
            rec[f"append_frame_seed_{a}_{ci}"] = frame_seed
            start = (start + n) % actor_capacity
            if ci == 1:
                fr.set("is_full_actor:" + str(a), 1)
                tree.is_full_actor[a] = 1
        heads.append(start)
    assert np.array_equal(_ref_tree_array(fr, cap), tree.tree), "tree mismatch after appends"
    rec["tree_after_append"] = tree.tree.copy()
    rec["heads"] = np.array(heads)
    for r in range(rounds):
        p_total = mem.transitions.total()
        u = rs.uniform(size=batch)
        perm = rs.permutation(batch)
        samples = osum.stratified_samples(p_total, batch, u, perm)
        # reference: descend with the injected samples, then the rest of find_multiple_values by hand-off
        ref_idx = mem.transitions._retrieve_multiple_values(np.zeros(batch, dtype=int), samples.copy())
        tab_index_actor = np.array([int(fr.get("index_actor:" + str(a))) for a in range(nb_actor)])
        ref_valid = mem.transitions.transform_to_valid_tree_indexes(ref_idx.copy(), tab_index_actor, 4, 3)
        ref_pri = np.array([float(fr.get("priorities:" + str(i))) for i in ref_valid])
        o_pri, o_data, o_idx, o_tot = tree.find(samples, 4, 3)
        assert np.array_equal(ref_valid, o_idx) and np.array_equal(ref_pri, o_pri) and o_tot == p_total
        capn = mem.transitions.get_current_capacity()
        ref_w = (capn * (ref_pri / p_total)) ** -mem.priority_weight
        ref_w = ref_w / ref_w.max()
        assert np.array_equal(ref_w, osum.importance_weights(o_pri, o_tot, tree.get_current_capacity(), 0.4))
        # transition assembly through the reference's byte path
        tab = mem.transitions.get_byte_multiple_transition(o_data, 4, 3)
        st, ac, rt, nx, nt = mem.get_torch_tensor_from_byte_transition(tab, batch)
        ost, oac, ort, onx, ont = store.assemble(o_data)
        assert np.array_equal((st * 255).round().numpy().astype(np.uint8), ost)
        assert torch.equal(st, torch.from_numpy(ost).float().div_(255))
        assert torch.equal(nx, torch.from_numpy(onx).float().div_(255))
        assert np.array_equal(ac.numpy(), oac) and np.array_equal(rt.numpy(), ort) and np.array_equal(nt.numpy(), ont)
        # priority update (with a forced duplicate to pin the double-count quirk)
        new_loss = rs.uniform(0.0, 2.0, batch).astype(np.float32)
        upd_idx = o_idx.copy()
        upd_idx[1] = upd_idx[0]
        mem.update_priorities(upd_idx, new_loss)
        tree.update_priorities(upd_idx, new_loss, 0.2)
        assert np.array_equal(_ref_tree_array(fr, cap), tree.tree), f"tree mismatch after update {r}"
        assert float(fr.get("max_priority")) == tree.max_priority
        rec[f"samples_{r}"] = samples
        rec[f"tree_idx_{r}"] = o_idx
        rec[f"pri_{r}"] = o_pri
        rec[f"weights_{r}"] = ref_w
        rec[f"p_total_{r}"] = p_total
        rec[f"upd_idx_{r}"] = upd_idx
        rec[f"upd_loss_{r}"] = new_loss
        rec[f"upd_pri_{r}"] = np.power(new_loss, 0.2)   # what redis_memory.py:560 produced on this host
        rec[f"tree_after_update_{r}"] = tree.tree.copy()
        rec[f"max_priority_{r}"] = tree.max_priority
        rec[f"asm_actions_{r}"] = oac
        rec[f"asm_returns_{r}"] = ort
        rec[f"asm_nonterminals_{r}"] = ont
        rec[f"asm_state_digest_{r}"] = np.array([int(ost.astype(np.int64).sum()), int(onx.astype(np.int64).sum()),
                                                 int((ost == 0).all(axis=(2, 3)).sum()), int((onx == 0).all(axis=(2, 3)).sum())])
    print(f"[{name}] tree/replay oracle == reference over {rounds} rounds; max parent-children error {tree.check():.2e}")
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)


def golden_actor(name, batch_size, len_buffer, cfg, seed):
    """Actor.act (actor.py:15-25) and Actor.compute_priorities (actor.py:41-124) of the unmodified reference on one
    synthetic actor buffer (an episode end in the middle), plus the launch_actor.py:127-133 tail rule."""
    from rainbowiqn.actor import Actor
    params = net.make_params(seed)
    inj = Injector()
    for _ in range(2):
        inj.push_noise(net.make_noise(99))
    with inj:
        actor = Actor(ref_args(batch_size, cfg), 18, None)
    actor.online_net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
    actor.update_target_net()
    actor.train()
    rs = np.random.RandomState(seed)
    n, hist = cfg["n_step"], 4
    frames = rs.randint(0, 256, (len_buffer + hist - 1, 84, 84)).astype(np.uint8)
    tab_state = [frames[i] for i in range(len(frames))]
    tab_action = [int(a) for a in rs.randint(0, 18, len_buffer)]
    tab_reward = [float(r) for r in rs.randint(-1, 2, len_buffer)]
    tab_nonterminal = [True] * len_buffer
    tab_nonterminal[len_buffer // 2] = False
    K, Np, N = cfg["n_quantile"], cfg["n_tau_prime"], cfg["n_tau"]
    ocfg = dict(cfg)
    # --- act: reset_noise (launch_actor.py:76-77) then the greedy action
    act_noise = net.make_noise(seed + 1)
    act_tau = rs.uniform(0, 1, (K, 1)).astype(np.float32)
    inj = Injector()
    inj.push_noise(act_noise)
    inj.push_tau(act_tau)
    with inj:
        actor.reset_noise()
        a_ref = actor.act(tab_state[:hist])
    p_on = net.apply_noise(net.to_torch(params), act_noise)
    a_or, q_mean = oactor.act(p_on, tab_state[:hist], K, torch.from_numpy(act_tau))
    assert a_ref == a_or, (a_ref, a_or)
    # --- compute_priorities
    n_tr = len_buffer - n
    chunks = [(lo, min(lo + batch_size, n_tr)) for lo in range(0, n_tr, batch_size)]
    noises = [cases.make_noises(seed + 100 + c) for c in range(len(chunks))]
    taus = [tuple(rs.uniform(0, 1, (nq * (hi - lo), 1)).astype(np.float32) for nq in (K, Np, N)) for lo, hi in chunks]
    inj = Injector()
    for c in range(len(chunks)):
        for k in range(3):
            inj.push_noise(noises[c][k])
        for t in taus[c]:
            inj.push_tau(t)
    # the reference interleaves reset_noise / forward per pass; the queues are consumed in the same relative order
    with inj:
        pri_ref = actor.compute_priorities(tab_state, tab_action, tab_reward, tab_nonterminal, 0.2)
    assert not inj.noise_q and not inj.tau_q
    pri_or = oactor.compute_priorities(net.to_torch(params), net.to_torch(params), tab_state, tab_action, tab_reward,
                                       tab_nonterminal, 0.2, noises,
                                       [tuple(torch.from_numpy(t) for t in tt) for tt in taus], ocfg, batch_size)
    err = float(np.max(np.abs(pri_ref - pri_or) / np.abs(pri_ref)))
    print(f"[{name}] act = {a_ref}; priorities max-rel-diff oracle vs reference = {err:.3e} over {n_tr} transitions")
    assert err < 1e-5, err
    rec = {"batch_size": batch_size, "len_buffer": len_buffer, "seed": seed, **{f"cfg_{k}": v for k, v in cfg.items()},
           "frames": frames, "tab_action": np.array(tab_action), "tab_reward": np.array(tab_reward),
           "tab_nonterminal": np.array(tab_nonterminal), "act_tau": act_tau, "act_action": a_ref,
           "act_q_mean": q_mean.numpy(), "priorities": pri_ref,
           "flushed": oactor.flush_priorities(pri_ref, 1.25, n)}
    for c, tt in enumerate(taus):
        for k, t in enumerate(tt):
            rec[f"tau_{c}_{k}"] = t
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)


def main():
    _install_stubs()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    golden_iqn("iqn_small", batch=4, cfg=cases.iqn_cfg(8, 8, 4), steps=2, seed=101)
    golden_iqn("iqn_cfg1", batch=32, cfg=cases.iqn_cfg(8, 8, 32), steps=1, seed=202)
    golden_c51("c51_small", batch=4, steps=2, seed=303)
    golden_sumtree("tree_pow2", actor_capacity=128, nb_actor=2, batch=32, rounds=3, seed=404)
    golden_sumtree("tree_npow2", actor_capacity=100, nb_actor=3, batch=40, rounds=3, seed=505)
    golden_actor("actor_small", batch_size=8, len_buffer=22, cfg=cases.iqn_cfg(8, 8, 4), seed=606)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
