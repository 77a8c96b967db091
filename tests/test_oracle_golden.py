"""CPU: the oracle restatement reproduces the fixtures recorded from the unmodified reference
(tests/golden/*.npz, written by oracle/make_golden.py in the dev container)."""
import os

import numpy as np
import torch

from oracle import actor as oactor, cases, losses, network as net, replay as oreplay, sumtree as osum


def _cfg(g):
    return cases.iqn_cfg(int(g["cfg_n_tau"]), int(g["cfg_n_tau_prime"]), int(g["cfg_n_quantile"]),
                         float(g["cfg_discount"]), int(g["cfg_n_step"]), float(g["cfg_kappa"]))


def _run_iqn(g):
    seed, batch, steps = int(g["seed"]), int(g["batch"]), int(g["steps"])
    cfg = _cfg(g)
    params = net.make_params(seed)
    p_on, p_tg = net.to_torch(params, requires_grad=True), net.to_torch(params)
    adam = losses.Adam([k for k in p_on if net.is_trainable(k)], lr=5e-5, eps=3.125e-4)
    for s in range(steps):
        b = cases.make_batch(seed + 10 + s, batch, n_step=cfg["n_step"], discount=cfg["discount"])
        taus = tuple(torch.from_numpy(t) for t in cases.make_taus(seed + 20 + s, batch, cfg))
        noises = cases.make_noises(seed + 30 + s)
        keep = {}
        loss, grads = losses.learn_step(p_on, p_tg, adam, cases.batch_to_torch(b), torch.from_numpy(b["weights"]),
                                        noises, taus, cfg, keep=keep)
        yield s, loss, grads, keep, p_on


def test_iqn_small_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "iqn_small.npz"))
    for s, loss, grads, keep, p_on in _run_iqn(g):
        assert np.allclose(loss.numpy(), g[f"loss_{s}"], rtol=1e-5, atol=0)
        assert np.array_equal(keep["a_star"].numpy(), g[f"a_star_{s}"])
        assert np.allclose(keep["theta"].detach().numpy(), g[f"theta_{s}"], rtol=1e-5, atol=1e-6)
        assert np.allclose(keep["target"].numpy(), g[f"target_{s}"], rtol=1e-5, atol=1e-6)
        for k, gr in grads.items():
            assert np.allclose(cases.tensor_digest(gr), g[f"grad_{s}_{k}"], rtol=2e-4, atol=1e-7), k
            assert np.allclose(cases.tensor_digest(p_on[k]), g[f"param_{s}_{k}"], rtol=1e-5, atol=1e-6), k


def test_iqn_cfg1_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "iqn_cfg1.npz"))
    for s, loss, grads, keep, p_on in _run_iqn(g):
        assert np.allclose(loss.numpy(), g[f"loss_{s}"], rtol=1e-5, atol=0)
        assert np.array_equal(keep["a_star"].numpy(), g[f"a_star_{s}"])


def test_c51_small_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "c51_small.npz"))
    seed, batch, steps = int(g["seed"]), int(g["batch"]), int(g["steps"])
    ocfg = dict(atoms=51, v_min=-10.0, v_max=10.0, discount=0.99, n_step=3)
    params = net.make_params(seed, rainbow_only=True)
    p_on, p_tg = net.to_torch(params, requires_grad=True), net.to_torch(params)
    adam = losses.Adam([k for k in p_on if net.is_trainable(k)], lr=6.25e-5, eps=1.5e-4)
    for s in range(steps):
        b = cases.make_batch(seed + 10 + s, batch)
        noises = cases.make_noises(seed + 30 + s, rainbow_only=True)
        keep = {}
        loss, grads = losses.learn_step(p_on, p_tg, adam, cases.batch_to_torch(b), torch.from_numpy(b["weights"]),
                                        noises, None, ocfg, rainbow_only=True, keep=keep)
        assert np.allclose(loss.numpy(), g[f"loss_{s}"], rtol=1e-5, atol=0)
        assert np.array_equal(keep["a_star"].numpy(), g[f"a_star_{s}"])
        assert np.allclose(keep["m"].numpy(), g[f"m_{s}"], rtol=1e-5, atol=1e-7)


def replay_tree_golden(g):
    """Re-run the recorded append sequence; yields (round, tree, store) before each sample/update round."""
    cap, nb, rounds = (int(g[k]) for k in ("actor_capacity", "nb_actor", "rounds"))
    tree, store = osum.SumTree(cap, nb), oreplay.ReplayStore(cap, nb)
    for a in range(nb):
        for ci in range(3):
            start, n = (int(v) for v in g[f"append_{a}_{ci}"])
            frames = np.random.RandomState(int(g[f"append_frame_seed_{a}_{ci}"])).randint(0, 256, (n, 84, 84)).astype(np.uint8)
            tree.append_priorities(start, a, g[f"append_pri_{a}_{ci}"])
            store.write(a, start, g[f"append_ts_{a}_{ci}"], frames, g[f"append_act_{a}_{ci}"],
                        g[f"append_rew_{a}_{ci}"], g[f"append_done_{a}_{ci}"])
            if ci == 1:
                tree.is_full_actor[a] = 1
    yield -1, tree, store
    for r in range(rounds):
        yield r, tree, store


def _check_tree_golden(g):
    for r, tree, store in replay_tree_golden(g):
        if r < 0:
            assert np.array_equal(tree.tree, g["tree_after_append"])
            assert np.array_equal(tree.index_actor, g["heads"])
            continue
        pri, data, idx, tot = tree.find(g[f"samples_{r}"], 4, 3)
        assert np.array_equal(idx, g[f"tree_idx_{r}"]) and np.array_equal(pri, g[f"pri_{r}"])
        assert tot == float(g[f"p_total_{r}"])
        w = osum.importance_weights(pri, tot, tree.get_current_capacity(), 0.4)
        assert np.array_equal(w, g[f"weights_{r}"])
        st, ac, rt, nx, nt = store.assemble(data)
        assert np.array_equal(ac, g[f"asm_actions_{r}"]) and np.array_equal(rt, g[f"asm_returns_{r}"])
        assert np.array_equal(nt, g[f"asm_nonterminals_{r}"])
        dg = np.array([int(st.astype(np.int64).sum()), int(nx.astype(np.int64).sum()),
                       int((st == 0).all(axis=(2, 3)).sum()), int((nx == 0).all(axis=(2, 3)).sum())])
        assert np.array_equal(dg, g[f"asm_state_digest_{r}"])
        tree.update_priorities(g[f"upd_idx_{r}"], g[f"upd_loss_{r}"], 0.2)
        assert np.array_equal(tree.tree, g[f"tree_after_update_{r}"])
        assert tree.max_priority == float(g[f"max_priority_{r}"])


def test_tree_pow2_matches_reference(golden_dir):
    _check_tree_golden(np.load(os.path.join(golden_dir, "tree_pow2.npz")))


def test_tree_npow2_matches_reference(golden_dir):
    _check_tree_golden(np.load(os.path.join(golden_dir, "tree_npow2.npz")))


def test_tree_edge_cases():
    # duplicated index double-counts (p1 + p2 - old), empty tree retrieves leaf 0 path, zero priorities
    t = osum.SumTree(8, 1)
    t.update_multiple_value(np.array([7, 8, 9]), np.array([1.0, 2.0, 3.0], np.float32))
    t.update_multiple_value(np.array([8, 8]), np.array([5.0, 7.0], np.float32))
    assert t.tree[8] == 5.0 + 7.0 - 2.0
    assert abs(t.tree[0] - (1.0 + 10.0 + 3.0)) < 1e-12
    idx = t.retrieve(np.array([0.0, 0.5, 1.0, 1.0001, 13.9999]))
    assert list(idx) == [7, 7, 7, 8, 9]


def actor_case(g):
    """Inputs of the actor fixture (shared with the GPU test): buffer lists, per-chunk noises / taus, cfg."""
    cfg = _cfg(g)
    seed, bs, lb = int(g["seed"]), int(g["batch_size"]), int(g["len_buffer"])
    frames = g["frames"]
    tab_state = [frames[i] for i in range(len(frames))]
    tab_action = [int(a) for a in g["tab_action"]]
    tab_reward = [float(r) for r in g["tab_reward"]]
    tab_nonterminal = [bool(x) for x in g["tab_nonterminal"]]
    n_chunks = -(-(lb - cfg["n_step"]) // bs)
    noises = [cases.make_noises(seed + 100 + c) for c in range(n_chunks)]
    taus = [tuple(torch.from_numpy(g[f"tau_{c}_{k}"]) for k in range(3)) for c in range(n_chunks)]
    return cfg, seed, bs, tab_state, tab_action, tab_reward, tab_nonterminal, noises, taus


def test_actor_small_matches_reference(golden_dir):
    """Actor.act + Actor.compute_priorities + the max_priority tail rule against the recorded reference outputs."""
    g = np.load(os.path.join(golden_dir, "actor_small.npz"))
    cfg, seed, bs, tab_state, tab_action, tab_reward, tab_nonterminal, noises, taus = actor_case(g)
    params = net.make_params(seed)
    p_on = net.apply_noise(net.to_torch(params), net.make_noise(seed + 1))
    a, q_mean = oactor.act(p_on, tab_state[:4], cfg["n_quantile"], torch.from_numpy(g["act_tau"]))
    assert a == int(g["act_action"]) and np.allclose(q_mean.numpy(), g["act_q_mean"], rtol=1e-5, atol=1e-6)
    pri = oactor.compute_priorities(net.to_torch(params), net.to_torch(params), tab_state, tab_action, tab_reward,
                                    tab_nonterminal, 0.2, noises, taus, cfg, bs)
    assert pri.shape == (int(g["len_buffer"]) - cfg["n_step"],)
    assert np.allclose(pri, g["priorities"], rtol=1e-5, atol=0)
    fl = oactor.flush_priorities(g["priorities"], 1.25, cfg["n_step"])
    assert np.array_equal(fl, g["flushed"]) and np.all(fl[-cfg["n_step"]:] == 1.25)
