"""Microbenchmark of riqn_quantile_embed_fwd_tc (the embedding producer): CUDA-event time per launch and achieved GB/s on the
algorithmic bytes of SURVEY 8d for the three launch shapes of a learner step.
    python tools/time_embed.py            # fp16 mode: K=32 one image, N'=64 one image, N=64 two images; bf16x3 mode: hi + lo"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rainbow_iqn_apex_b200._lib import call, ptr  # noqa: E402

dev = torch.device("cuda")
B, E, F = 512, 64, 3136


def bf(*s):
    return torch.empty(*s, device=dev, dtype=torch.bfloat16)


def run(nq, images, fp16, reps=20):
    R = B * nq
    tau = torch.rand(R, device=dev)
    feat = torch.rand(B, F, device=dev)
    w = torch.randn(F, E, device=dev) * 0.1
    w_hi = w.to(torch.bfloat16)
    w_lo = (w - w_hi.float()).to(torch.bfloat16)
    bias = torch.randn(F, device=dev) * 0.1
    cos_hi, cos_lo = bf(R, E), bf(R, E)
    sets = [(bf(R, F), bf(R, F) if images == 2 else None) for _ in range(2)]       # rotate outputs (> L2)

    def go(i):
        x_hi, x_lo = sets[i & 1]
        call("riqn_quantile_embed_fwd_tc", B, nq, E, F, ptr(tau), ptr(feat), ptr(w_hi), ptr(w_lo), ptr(bias), ptr(cos_hi),
             ptr(cos_lo), None, None, ptr(x_hi), ptr(x_lo), None, None, 1 if fp16 else 0)
    for i in range(3):
        go(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        go(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    nbytes = 2.0 * images * R * F + 4.0 * R + 4.0 * B * F + 4.0 * (E * F + F)
    print(f"nq={nq:3d} images={images} {'fp16' if fp16 else 'bf16 hi/lo'}: {us:8.1f} us/launch (cos kernel included)  "
          f"{nbytes / us / 1e3:6.0f} GB/s of {nbytes / 1e6:.0f} MB algorithmic")


if __name__ == "__main__":
    run(32, 1, True)
    run(64, 1, True)
    run(64, 2, True)
    run(64, 2, False)
