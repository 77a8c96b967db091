"""CPU, world_size = 2 over gloo: host logic of the data-parallel learner (rainbow_iqn_apex_b200/parallel.py).
No kernel is launched: the networks stay on the CPU (their arenas are ordinary tensors there); what is checked is the
replica bootstrap (parameters / epsilons / noise seeds broadcast from rank 0, rank-private quantile streams) and the
gradient reduction contract (sum over ranks of the flat arena, grad_scale = 1/world for the Adam kernel)."""
import os
import socket
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import make_args


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from rainbow_iqn_apex_b200 import parallel
    from rainbow_iqn_apex_b200.model import DQN
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                      # different initial weights / seeds per rank
    args = make_args(torch.device("cpu"))
    learner = SimpleNamespace(online_net=DQN(args, 18), target_net=DQN(args, 18),
                              optimiser=SimpleNamespace(grad_scale=1.0), process_group=None)
    before = learner.online_net._flat.clone()
    parallel.make_data_parallel(learner)
    on = learner.online_net
    # 1. replicas identical after bootstrap, seeds shared, tau streams private
    flat0 = on._flat.clone()
    dist.broadcast(flat0, src=0)
    assert torch.equal(flat0, on._flat)
    seeds = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(seeds, torch.tensor([on._rng_seed]))
    assert all(int(s) == int(seeds[0]) for s in seeds)
    assert on._tau_stream_offset == rank << 40
    assert learner.optimiser.grad_scale == 1.0 / world and learner.process_group is not None
    if rank == 1:
        assert not torch.equal(before, on._flat)       # rank 1 really received rank 0's parameters
    # 2. gradient reduction: the arena all-reduce yields the sum, and the views in .grad see it
    on._flat_grad.fill_(float(rank + 1))
    dist.all_reduce(on._flat_grad, group=learner.process_group)
    assert float(on.conv1.weight.grad.flatten()[0]) == sum(range(1, world + 1))
    # 3. Ape-X parameter publication: rank 0 (learner) -> everyone (actors)
    agent = SimpleNamespace(online_net=on)
    if rank == 0:
        on._flat.add_(1.0)
    parallel.publish_parameters(agent, src=0)
    ref = on._flat.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, on._flat)
    assert parallel.shard_batch(4096, world) == 4096 // world
    assert parallel.allreduce_max(float(rank), torch.device("cpu")) == world - 1
    out.put((rank, "ok"))
    dist.destroy_process_group()


def test_data_parallel_bootstrap_and_grad_reduction_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(world))
    assert got == [(0, "ok"), (1, "ok")]
