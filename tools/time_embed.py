"""Microbenchmark of riqn_quantile_embed_fwd_tc with different output sets (which images are written)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rainbow_iqn_apex_b200._lib import call, ptr
dev = torch.device("cuda")
B, E, F = 512, 64, 3136
def bf(*s): return torch.empty(*s, device=dev, dtype=torch.bfloat16)
def run(nq, outs, reps=10):
    R = B * nq
    tau = torch.rand(R, device=dev); feat = torch.rand(B, F, device=dev)
    w = torch.randn(F, E, device=dev) * 0.1
    w_hi = w.to(torch.bfloat16); w_lo = (w - w_hi.float()).to(torch.bfloat16)
    bias = torch.randn(F, device=dev) * 0.1
    cos_hi, cos_lo, cosT = bf(R, E), bf(R, E), bf(E, R)
    x32 = torch.empty(R, F, device=dev) if "x32" in outs else None
    x_hi = bf(R, F) if "hi" in outs else None
    x_lo = bf(R, F) if "lo" in outs else None
    x_hiT = bf(F, R) if "hiT" in outs else None
    x_loT = bf(F, R) if "loT" in outs else None
    def go():
        call("riqn_quantile_embed_fwd_tc", B, nq, E, F, ptr(tau), ptr(feat), ptr(w_hi), ptr(w_lo), ptr(bias), ptr(cos_hi),
             ptr(cos_lo), ptr(cosT), ptr(x32), ptr(x_hi), ptr(x_lo), ptr(x_hiT), ptr(x_loT))
    for _ in range(3): go()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    nbytes = R * F * sum({"x32": 4, "hi": 2, "lo": 2, "hiT": 2, "loT": 2}[o] for o in outs)
    print(f"nq={nq} outs={','.join(outs):20s}: {us:8.1f} us  out {nbytes/us/1e3:6.0f} GB/s")
only = sys.argv[1:] 
if only:
    run(64, only[0].split(","), reps=2)
else:
    run(64, ["hi"]); run(64, ["hi", "lo"]); run(64, ["hi", "lo", "hiT"]); run(64, ["x32"]); run(32, ["hi", "lo"])
