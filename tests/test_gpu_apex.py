"""GPU (one device): the Ape-X topology's device halves (rainbow_iqn_apex_b200/apex.py; BASELINE configs[3]) -- shard
sampling + batch assembly + priority routing against the numpy oracle, and the batched actor pool (act / observe / flush
with initial priorities and the max_priority tail rule) against Actor.compute_priorities, which itself is pinned on the
reference fixture (tests/test_gpu_parity_full.py::test_actor_matches_reference_golden)."""
import numpy as np
import pytest
import torch

from helpers import load_params, make_args
from oracle import cases, network as net, sumtree as osum

pytestmark = pytest.mark.gpu


def _shard(dev, cap, nb, seed, batch=32):
    from rainbow_iqn_apex_b200 import ReplayMemory
    rs = np.random.RandomState(seed)
    mem = ReplayMemory(make_args(dev, batch, nb_actor=nb, actor_capacity=cap), None)
    ot = osum.SumTree(cap, nb)
    for a in range(nb):
        pri = (rs.uniform(0.05, 1, cap) ** 0.2).astype(np.float32)
        frames = rs.randint(0, 256, (cap, 84, 84)).astype(np.uint8)
        mem.transitions.append_arrays(a, 0, np.arange(cap) % 23, frames, rs.randint(0, 18, cap),
                                      rs.randint(-1, 2, cap).astype(np.float32), rs.uniform(size=cap) < 0.04, pri)
        ot.append_priorities(0, a, pri)
        ot.is_full_actor[a] = 1
    return mem, ot, rs


def test_sharded_sample_assemble_route_vs_oracle(cuda_dev):
    from rainbow_iqn_apex_b200 import apex
    S, B = 3, 20
    counts = apex.shard_counts(B, S)
    assert counts == [7, 7, 6]
    shards = [_shard(cuda_dev, 96, 2, 40 + s) for s in range(S)]
    parts, o_pri, o_idx = [], [], []
    for (mem, ot, rs), c in zip(shards, counts):
        samples = osum.stratified_samples(ot.total(), c, rs.uniform(size=c), rs.permutation(c))
        parts.append(apex.sample_shard(mem, c, max(counts), samples=samples))
        p, d, i, _ = ot.find(samples, 4, 3)
        o_pri.append(p)
        o_idx.append(i)
        assert np.array_equal(parts[-1]["tree_idx"][:c].cpu().numpy(), i)          # bit-exact descent per shard
        assert np.array_equal(parts[-1]["pri"][:c].cpu().numpy(), p)
        assert parts[-1]["window"].shape == (max(counts), 7, 84, 84)
    totals = torch.stack([m.transitions.tree[0] for m, _, _ in shards])
    filled = float(sum(m.transitions.get_current_capacity() for m, _, _ in shards))
    shard_of, tree_idx, st, ac, rt, nx, nt, w = apex.assemble_batch(parts, counts, totals, filled, 0.4)
    assert st.shape == (B, 4, 84, 84) and nx.shape == (B, 4, 84, 84) and torch.equal(st[:, 3], nx[:, 0])
    pri = np.concatenate(o_pri)
    tot = np.array([shards[s][1].total() for s in shard_of.tolist()])
    cnt = np.array([counts[s] for s in shard_of.tolist()], np.float64)
    wn = (filled * (cnt / B) * pri / tot) ** -0.4
    assert np.allclose(w.cpu().numpy(), (wn / wn.max()).astype(np.float32), rtol=1e-6)
    # route a loss vector back: every shard's tree == the oracle tree updated with its own slice
    loss = torch.from_numpy(np.random.RandomState(1).uniform(0.1, 2, B).astype(np.float32)).to(cuda_dev)
    lo = 0
    for s, ((mem, ot, _), c) in enumerate(zip(shards, counts)):
        new_pri = apex.route_priorities(mem, s, counts, parts[s], loss)
        ot.update_multiple_value(o_idx[s], new_pri.cpu().numpy())
        assert np.array_equal(mem.transitions.tree.cpu().numpy(), ot.tree)
        lo += c


def test_actor_pool_flush_matches_compute_priorities(cuda_dev):
    """One environment, one buffer: ActorPool.flush == Actor.compute_priorities + tail rule + append (same injected
    randomness -> same kernels on the same inputs), and the shard holds exactly the buffered steps afterwards."""
    from rainbow_iqn_apex_b200 import Actor, ReplayMemory, apex
    cfg = cases.iqn_cfg(8, 8, 4)
    bs, L, seed = 8, 22, 515
    args = make_args(cuda_dev, bs, cfg, nb_actor=1, actor_capacity=64)
    actor = Actor(args, 18, None)
    load_params(actor.online_net, net.make_params(seed))
    actor.update_target_net()
    actor.train()
    mem = ReplayMemory(args, None)
    mem.transitions.max_priority.fill_(1.5)
    pool = apex.ActorPool(actor, mem, 1, L)
    rs = np.random.RandomState(seed)
    frames = rs.randint(0, 256, (L + 3, 84, 84)).astype(np.uint8)
    acts = rs.randint(0, 18, L)
    rews = rs.randint(-1, 2, L).astype(np.float32)
    dones = np.zeros(L, bool)
    dones[9] = True
    dev = cuda_dev
    for i in range(L):
        stack = torch.from_numpy(frames[i:i + 4])[None].to(dev)
        full = pool.observe(stack, torch.tensor([acts[i]], device=dev), torch.tensor([rews[i]], device=dev),
                            torch.tensor([bool(dones[i])], device=dev))
        assert full == (i == L - 1)
    n_chunks = -(-(L - 3) // bs)
    inj = [dict(noises=cases.make_noises(seed + 100 + c),
                taus=tuple(torch.from_numpy(t) for t in cases.make_taus(seed + 200 + c, min(bs, L - 3 - c * bs), cfg)))
           for c in range(n_chunks)]
    actor._inject = [dict(d) for d in inj]
    pri_pool = pool.initial_priorities().cpu().numpy()[0]
    actor._inject = [dict(d) for d in inj]
    pri_ref = actor.compute_priorities([frames[i] for i in range(L + 3)], [int(a) for a in acts], [float(r) for r in rews],
                                       [not d for d in dones], 0.2)
    assert pri_pool.shape == pri_ref.shape == (L - 3,)
    assert np.allclose(pri_pool, pri_ref, rtol=2e-6, atol=0)           # device powf vs numpy's float32 power
    actor._inject = [dict(d) for d in inj]
    assert pool.flush(T_actor=L) == L and pool.fill == 0 and pool.write_index[0] == L
    tr = mem.transitions
    C = tr.full_capacity
    leaves = tr.tree[C - 1:C - 1 + L].cpu().numpy()
    assert np.all(leaves[-3:] == 1.5)                                  # launch_actor.py:127-133
    assert np.allclose(leaves[:-3], pri_ref, rtol=2e-6)
    assert np.array_equal(tr.frames[:L].cpu().numpy().reshape(L, 84, 84), frames[3:])
    assert np.array_equal(tr.action[:L].cpu().numpy(), acts) and np.array_equal(tr.reward[:L].cpu().numpy(), rews)
    assert np.array_equal(tr.nonterminal[:L].cpu().numpy().astype(bool), ~dones)
    ts = tr.timestep[:L].cpu().numpy()
    assert ts[0] == 0 and ts[9] == 9 and ts[10] == 0 and ts[11] == 1   # the episode counter restarts after a terminal step
    assert tr.check_sumtree_correct() < 1e-12 and tr.get_current_capacity() == L


def test_actor_pool_many_envs(cuda_dev):
    """E = 6 environments, two flushes with native device randomness: segments advance independently, the tree stays
    consistent, sampling from the filled part works and act() returns one action per environment."""
    from rainbow_iqn_apex_b200 import Actor, ReplayMemory, apex
    cfg = cases.iqn_cfg(8, 8, 4)
    E, L = 6, 10
    args = make_args(cuda_dev, 16, cfg, nb_actor=E, actor_capacity=32)
    actor = Actor(args, 18, None)
    actor.train()
    mem = ReplayMemory(args, None)
    pool = apex.ActorPool(actor, mem, E, L)
    g = torch.Generator(device=cuda_dev).manual_seed(3)
    states = torch.randint(0, 256, (E, 4, 84, 84), dtype=torch.uint8, device=cuda_dev, generator=g)
    appended = 0
    for step in range(2 * L):
        a = pool.act(states)
        assert a.shape == (E,) and int(a.min()) >= 0 and int(a.max()) < 18
        nxt = torch.randint(0, 256, (E, 1, 84, 84), dtype=torch.uint8, device=cuda_dev, generator=g)
        rew = (torch.randint(0, 3, (E,), device=cuda_dev, generator=g) - 1).float()
        done = torch.rand(E, device=cuda_dev, generator=g) < 0.1
        if pool.observe(states, a, rew, done):
            appended += pool.flush(T_actor=step)
        states = torch.cat([states[:, 1:], nxt], 1)
    assert appended == 2 * L * E and list(pool.write_index) == [2 * L] * E
    tr = mem.transitions
    assert tr.get_current_capacity() == 2 * L * E and tr.check_sumtree_correct() < 1e-9
    assert float(tr.tree[0]) > 0 and float(tr.max_priority) >= 1.0
    idx, st, ac, rt, nx, nt, w = mem.sample(16)
    assert torch.isfinite(w).all() and st.shape == (16, 4, 84, 84)
