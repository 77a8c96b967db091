import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from helpers import make_args
from oracle import sumtree as osum
from rainbow_iqn_apex_b200 import ReplayMemory
dev = torch.device("cuda")
np.set_printoptions(precision=17, linewidth=200)
g = np.load(os.path.join(R, "tests/golden/tree_pow2.npz"))
cap, nb = int(g["actor_capacity"]), int(g["nb_actor"])
mem = ReplayMemory(make_args(dev, 32, nb_actor=nb, actor_capacity=cap), None); tr = mem.transitions
ot = osum.SumTree(cap, nb)
C = cap * nb
for a in range(nb):
    for ci in range(3):
        start, n = (int(v) for v in g[f"append_{a}_{ci}"])
        pri = g[f"append_pri_{a}_{ci}"]
        pos = (np.arange(start, start + n) % cap) + a * cap
        idx = pos + C - 1
        before = tr.tree.cpu().numpy().copy()
        tr.update_multiple_value(torch.from_numpy(idx).to(dev), torch.from_numpy(pri).to(dev))
        ot.update_multiple_value(idx, pri)
        got = tr.tree.cpu().numpy()
        bad = np.where(got != ot.tree)[0]
        print("actor", a, "chunk", ci, "start", start, "n", n, "nbad", len(bad), "first bad", bad[:10])
        if len(bad):
            i = bad[0]
            print("  node", i, "before", before[i], "got", got[i], "ref", ot.tree[i], "got-before", got[i]-before[i], "ref-before", ot.tree[i]-before[i])
            # which batch entries touch node i
            lvl_nodes = idx.copy(); contrib = []
            while lvl_nodes.max() > 0:
                contrib += [j for j in range(n) if lvl_nodes[j] == i]
                lvl_nodes = (lvl_nodes - 1) // 2
            print("  entries touching it", contrib[:20], "n", len(contrib))
            tr.tree.copy_(torch.from_numpy(ot.tree))   # resync
