"""CPU: the host logic of the strip convolution (model._strip_perm, block grids, row shifts) -- the arithmetic the CUDA
path implements with TMA tiles (csrc/gemm_tc.cu TC_CONV, csrc/conv.cu) -- against torch's conv2d.  Reference layers:
rainbowiqn/model.py:65-67."""
import pytest
import torch
import torch.nn.functional as F

from rainbow_iqn_apex_b200.model import _strip_perm

LAYERS = [(4, 84, 32, 8, 4, 1, True), (32, 20, 64, 4, 2, 0, False), (64, 9, 64, 3, 1, 0, False)]


def _block_matrix(x, k, s, pad, first):
    """(B*G*G, s*s*C) block matrix of the zero-padded input; within-block order (c, iy, ix) for the first layer
    (riqn_s2d_u8), (iy, ix, c) for the others (written by the previous layer's epilogue)."""
    B, C, H, _ = x.shape
    t = k // s
    OH = (H + 2 * pad - k) // s + 1
    G = OH + t - 1
    xp = torch.zeros(B, C, G * s, G * s, dtype=x.dtype)
    hh = min(H, G * s - pad)
    xp[:, :, pad:pad + hh, pad:pad + hh] = x[:, :, :hh, :hh]
    blk = xp.view(B, C, G, s, G, s)                                  # b c gy iy gx ix
    A = blk.permute(0, 2, 4, 1, 3, 5) if first else blk.permute(0, 2, 4, 3, 5, 1)
    return A.reshape(B * G * G, C * s * s), G, OH, t


@pytest.mark.parametrize("layer", LAYERS)
def test_strip_formulation_equals_conv2d(layer):
    C, H, Co, k, s, pad, first = layer
    g = torch.Generator().manual_seed(C + k)
    x = torch.randn(3, C, H, H, generator=g, dtype=torch.float64)
    w = torch.randn(Co, C, k, k, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=s, padding=pad)
    A, G, OH, t = _block_matrix(x, k, s, pad, first)
    Kc = C * s * s
    perm = _strip_perm(C, k, s, first)
    assert sorted(perm.tolist()) == list(range(C * k * k))            # a permutation of the (c, kh, kw) index
    wp = w.reshape(Co, -1)[:, perm]
    Mp = A.shape[0]
    Apad = torch.cat([A, torch.zeros(t * G + t, Kc, dtype=A.dtype)])  # rows past the end: TMA zero fill
    out = torch.zeros(Mp, Co, dtype=torch.float64)
    for dy in range(t):
        for dx in range(t):
            sft = dy * t + dx                                         # k-block group `sft` reads rows m + dy*G + dx
            out += Apad[dy * G + dx: dy * G + dx + Mp] @ wp[:, sft * Kc:(sft + 1) * Kc].T
    got = out.view(3, G, G, Co)[:, :OH, :OH].permute(0, 3, 1, 2)      # only gy < OH, gx < OW are real outputs
    assert torch.allclose(got, ref, rtol=1e-10, atol=1e-10)
    # weight gradient in the strip order maps back through the same permutation (unpermute_add_kernel)
    dy_grid = torch.zeros(3, G, G, Co, dtype=torch.float64)
    dout = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    dy_grid[:, :OH, :OH] = dout.permute(0, 2, 3, 1)
    dYg = dy_grid.reshape(Mp, Co)
    dwp = torch.cat([dYg.T @ Apad[(sft // t) * G + sft % t: (sft // t) * G + sft % t + Mp] for sft in range(t * t)], dim=1)
    dw = torch.zeros(Co, C * k * k, dtype=torch.float64)
    dw[:, perm] += dwp
    wr = w.clone().requires_grad_(True)
    F.conv2d(x, wr, stride=s, padding=pad).backward(dout)
    assert torch.allclose(dw, wr.grad.reshape(Co, -1), rtol=1e-9, atol=1e-9)
