"""CPU: the ordering argument behind update_propagate_kernel (csrc/sumtree.cu), checked against the oracle's literal
restatement of the reference loop (oracle/sumtree.py: propagate, redis_memory.py:94-105).

The reference walks the batch level by level: at every level each entry adds its diff to its current ancestor, in batch
order.  The kernel instead lets ONE warp per (depth, node) replay that node's adds: the hits of the node sorted by the
number of parent steps between the leaf and the node (shallower leaves of a non-power-of-two tree reach a node on an
earlier level), and in batch order within one step count; the root gets numpy's pairwise sum.  The two orders are the
same sequence of float64 adds per node -- this test states that with numpy, duplicates and two-depth trees included."""
import numpy as np
import pytest

from oracle import sumtree as osum


def _per_node_replay(tree, idx, diff):
    """What the kernel does, node by node (plain Python; no warp mechanics)."""
    tree = tree.copy()
    depth = np.floor(np.log2(idx + 1)).astype(np.int64)
    for d in range(1, int(depth.max()) + 1):
        st = depth - d
        node = np.where(st >= 0, ((idx + 1) >> np.maximum(st, 0)) - 1, -1)
        for me in dict.fromkeys(node[node > 0].tolist()):          # first-occurrence order, any order would do
            hits = np.flatnonzero(node == me)
            acc = tree[me]
            for s in range(int(st[hits].min()), int(st[hits].max()) + 1):
                for k in hits[st[hits] == s]:
                    acc = acc + diff[k]
            tree[me] = acc
    tree[0] = tree[0] + (0.0 + np.sum(diff))
    return tree


@pytest.mark.parametrize("cap,nb,n", [(64, 1, 40), (37, 1, 25), (1000, 3, 512), (500, 2, 700), (1 << 12, 1, 512)])
def test_per_node_replay_equals_level_walk(cap, nb, n):
    rs = np.random.RandomState(cap * 7 + n)
    ot = osum.SumTree(cap, nb)
    C = cap * nb
    ot.update_multiple_value(np.arange(C) + C - 1, (rs.uniform(1e-3, 1, C) ** 0.2).astype(np.float32))
    for _ in range(3):
        idx = rs.randint(0, C, n).astype(np.int64) + C - 1
        idx[n // 2] = idx[n // 3] = idx[0]                           # triple duplicate
        pri = (rs.uniform(0, 3, n) ** 0.2).astype(np.float32)
        diff = pri.astype(np.float64) - ot.tree[idx]
        model = _per_node_replay(ot.tree, idx, diff)
        ot.update_multiple_value(idx, pri)
        assert np.array_equal(model, ot.tree)
