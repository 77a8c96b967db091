"""CPU, world_size = 3 over gloo (1 learner rank + 2 replay-shard ranks): the host logic of the Ape-X topology
(rainbow_iqn_apex_b200/apex.py; BASELINE configs[3]).  No kernel is launched: every shard is the numpy oracle
(oracle.sumtree / oracle.replay) behind the few methods apex.py calls on a ReplayMemory, so what is checked is the
routing arithmetic -- per-shard draw counts, the gather of the windows to the learner, importance weights against the
shard totals, the broadcast of the new losses back to the owning shard's leaves."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import replay as oreplay, sumtree as osum

CAP, NB, BATCH = 64, 2, 11          # per shard: 2 segments of 64 slots; 11 = 6 + 5 draws per learner step


class OracleShard:
    """The subset of ReplayMemory that apex.sample_shard / route_priorities use, on CPU tensors."""

    def __init__(self, seed):
        rs = np.random.RandomState(seed)
        self.device, self.history, self.n, self.priority_exponent = torch.device("cpu"), 4, 3, 0.2
        self.tree_o = osum.SumTree(CAP, NB)
        self.store = oreplay.ReplayStore(CAP, NB)
        for a in range(NB):
            pri = (rs.uniform(0.05, 1, CAP) ** 0.2).astype(np.float32)
            self.tree_o.append_priorities(0, a, pri)
            ts = np.arange(CAP) % 17
            self.store.write(a, 0, ts, rs.randint(0, 256, (CAP, 84, 84)).astype(np.uint8), rs.randint(0, 18, CAP),
                             rs.randint(-1, 2, CAP).astype(np.float64), rs.uniform(size=CAP) < 0.05)
            self.tree_o.index_actor[a] = int(rs.randint(0, CAP))
            self.tree_o.is_full_actor[a] = 1
        self.rs = rs
        shard = self

        class _Tr:
            @property
            def tree(self):
                return torch.from_numpy(shard.tree_o.tree)

            def get_current_capacity(self):
                return shard.tree_o.get_current_capacity()

            def find_multiple_values(self, history, n, count, samples=None):
                if samples is None:
                    samples = osum.stratified_samples(shard.tree_o.total(), count, shard.rs.uniform(size=count),
                                                      shard.rs.permutation(count))
                pri, data, idx, _ = shard.tree_o.find(np.asarray(samples), history, n)
                return torch.from_numpy(pri), torch.from_numpy(data), torch.from_numpy(idx)

        self.transitions = _Tr()

    def assemble_window(self, data_idx):
        st, ac, rt, nx, nt = self.store.assemble(data_idx.numpy())
        win = np.concatenate([st, nx[:, 1:]], 1)            # frames 0..6: states = 0:4, next_states = 3:7
        return torch.from_numpy(win), torch.from_numpy(ac), torch.from_numpy(rt), torch.from_numpy(nt)

    def update_priorities(self, idxs, loss):
        self.tree_o.update_priorities(idxs.numpy(), loss.numpy().astype(np.float32), self.priority_exponent)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from rainbow_iqn_apex_b200 import apex, parallel
    parallel.init_from_env(backend="gloo")
    topo = apex.ApexTopology(BATCH, publish_every=2)
    assert topo.counts == [6, 5] and topo.n_max == 6 and topo.is_learner == (rank == 0)
    cpu = torch.device("cpu")
    shards = [OracleShard(700 + s) for s in range(world - 1)]          # every rank can rebuild every shard (same seeds)
    mem = None if rank == 0 else shards[rank - 1]
    agent = SimpleNamespace(online_net=SimpleNamespace(_flat=torch.full((8,), float(rank)), compose_weights=lambda: None))
    for step in range(2):
        got = topo.sample(mem, beta=0.4, device=cpu)
        if rank == 0:
            shard_of, tree_idx, states, actions, returns, next_states, nonterminals, w = got
            assert states.shape == (BATCH, 4, 84, 84) and next_states.shape == (BATCH, 4, 84, 84) and w.shape == (BATCH,)
            assert shard_of.tolist() == [0] * 6 + [1] * 5
            # replay the same draws locally (same RandomState streams) and compare everything the learner received
            exp = []
            for s, sh in enumerate(shards):
                exp.append(apex.sample_shard(sh, topo.counts[s], topo.n_max))
            ref = apex.assemble_batch(exp, topo.counts, torch.tensor([sh.tree_o.total() for sh in shards], dtype=torch.float64),
                                      float(sum(sh.tree_o.get_current_capacity() for sh in shards)), 0.4)
            for i, (a, b) in enumerate(zip(got, ref)):
                assert torch.equal(a, b), (step, i)
            # weights: w_i = (N * (c_s / B) * p_i / total_s)^-beta / max, in float64
            pri = torch.cat([e["pri"][:c] for e, c in zip(exp, topo.counts)]).numpy()
            tot = np.array([shards[s].tree_o.total() for s in shard_of.tolist()])
            cnt = np.array([topo.counts[s] for s in shard_of.tolist()], np.float64)
            wn = (2 * CAP * NB * (cnt / BATCH) * pri / tot) ** -0.4
            assert np.allclose(w.numpy(), (wn / wn.max()).astype(np.float32), rtol=1e-6)
            loss = torch.from_numpy(np.random.RandomState(step).uniform(0.1, 2, BATCH).astype(np.float32))
            # the learner's own copy of the shards follows the same updates, to check the actor ranks' trees below
            for s, sh in enumerate(shards):
                apex.route_priorities(sh, s, topo.counts, exp[s], loss)
        else:
            loss = torch.empty(BATCH, dtype=torch.float32)
        topo.route(loss, mem, None if rank == 0 else got)
        published = topo.maybe_publish(agent)
        assert published == (step == 1)
    if rank != 0:
        out.put((rank, mem.tree_o.tree.copy(), float(agent.online_net._flat[0])))
    else:
        out.put((0, [sh.tree_o.tree.copy() for sh in shards], float(agent.online_net._flat[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_apex_topology_sampling_and_routing_gloo():
    world = 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = out.get(timeout=180)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    learner_trees = res[0][0]
    for rank in (1, 2):
        # the actor rank's tree after two routed updates == the learner-side replica that applied the same slices
        assert np.array_equal(res[rank][0], learner_trees[rank - 1])
        assert res[rank][1] == 0.0                                    # parameters arrived from the learner (rank 0)


def test_shard_counts_and_single_shard_weights():
    from rainbow_iqn_apex_b200 import apex
    assert apex.shard_counts(512, 7) == [74, 73, 73, 73, 73, 73, 73] and sum(apex.shard_counts(512, 7)) == 512
    assert apex.shard_counts(512, 1) == [512]
    rs = np.random.RandomState(3)
    pri = rs.uniform(0.01, 1, 40)
    w = apex.sharded_is_weights(torch.from_numpy(pri), torch.zeros(40, dtype=torch.int64), [pri.sum() * 3], [40], 5000, 0.4)
    # one shard: the reference's formula (redis_memory.py:465-475)
    assert np.allclose(w.numpy(), osum.importance_weights(pri, pri.sum() * 3, 5000, 0.4), rtol=1e-12)
    # a non-positive priority takes the 1/capacity fallback (redis_memory.py:446-456) and then carries the largest weight
    pri[5] = 0.0
    w = apex.sharded_is_weights(torch.from_numpy(pri), torch.zeros(40, dtype=torch.int64), [pri.sum()], [40], 5000, 0.4)
    assert w[5] == 1.0 and torch.isfinite(w).all()


def test_pack_unpack_roundtrip():
    """The packed per-shard record (one collective per learner batch) returns every field bit for bit, as views."""
    from rainbow_iqn_apex_b200 import apex
    n_max, g = 5, torch.Generator().manual_seed(4)
    smp = apex.ShardSample.empty(n_max, torch.device("cpu"))
    smp["tree_idx"].copy_(torch.randint(0, 1 << 40, (n_max,), generator=g))
    smp["pri"].copy_(torch.rand(n_max, dtype=torch.float64, generator=g))
    smp["window"].copy_(torch.randint(0, 256, (n_max, 7, 84, 84), dtype=torch.uint8, generator=g))
    smp["actions"].copy_(torch.randint(0, 18, (n_max,), generator=g))
    smp["returns"].copy_(torch.randn(n_max, generator=g))
    smp["nonterminals"].copy_((torch.rand(n_max, generator=g) < 0.9).float())
    stat = torch.tensor([123.456, 7890.0], dtype=torch.float64)
    buf = apex.pack(smp, stat, n_max)
    assert buf.dtype == torch.uint8 and buf.numel() == apex.packed_bytes(n_max)
    out, st = apex.unpack(buf, n_max)
    for k in apex.ShardSample.FIELDS:
        assert out[k].dtype == smp[k].dtype and torch.equal(out[k], smp[k]), k
    assert torch.equal(st, stat)
    # the filled capacity may arrive as a device scalar (no host sync on the learner): same weights as with a float
    pri, sh = torch.rand(8, dtype=torch.float64, generator=g) + 0.01, torch.tensor([0, 0, 0, 1, 1, 1, 1, 0])
    w1 = apex.sharded_is_weights(pri, sh, [3.0, 5.0], [4, 4], 1000.0, 0.4)
    w2 = apex.sharded_is_weights(pri, sh, torch.tensor([3.0, 5.0], dtype=torch.float64), [4, 4], torch.tensor(1000.0, dtype=torch.float64), 0.4)
    assert torch.equal(w1, w2)
