"""Microbenchmark of riqn_sumtree_update (priority write-back of one learner batch)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rainbow_iqn_apex_b200._lib import call, ptr
dev = torch.device("cuda")
def run(n, cap, reps=20):
    tree = torch.rand(2 * cap - 1, device=dev, dtype=torch.float64)
    idx = (torch.randperm(cap, device=dev)[:n] + cap - 1).long()
    loss = torch.rand(n, device=dev)
    newp = torch.empty(n, device=dev); diff = torch.empty(n, device=dev, dtype=torch.float64)
    mx = torch.ones(1, device=dev, dtype=torch.float64)
    go = lambda: call("riqn_sumtree_update", n, cap, ptr(tree), ptr(idx), ptr(loss), 0.5, 1, ptr(newp), ptr(diff), ptr(mx))
    for _ in range(3): go()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    print(f"n={n} capacity={cap}: {e0.elapsed_time(e1) * 1e3 / reps:8.1f} us")
if len(sys.argv) > 1:
    run(512, 1 << 19, reps=2)
else:
    run(512, 1 << 19); run(512, 1 << 14); run(64, 1 << 19); run(4096, 1 << 19)
