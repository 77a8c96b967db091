"""GPU: the tcgen05/TMA GEMM (csrc/gemm_tc.cu) against float64 products computed on the host."""
import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _bf16_round(x):
    return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()


def _run(dev, M, N, K, split3, epi=0, split_k=1, seed=0):
    from rainbow_iqn_apex_b200._lib import call, ptr
    rs = np.random.RandomState(seed)
    A = rs.standard_normal((M, K)).astype(np.float32)
    B = (rs.standard_normal((N, K)) * 0.05).astype(np.float32)
    a, b = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    a_hi = torch.empty(M, K, dtype=torch.bfloat16, device=dev)
    b_hi = torch.empty(N, K, dtype=torch.bfloat16, device=dev)
    a_lo = torch.empty_like(a_hi) if split3 else None
    b_lo = torch.empty_like(b_hi) if split3 else None
    call("riqn_split_bf16", M, K, ptr(a), ptr(a_hi), ptr(a_lo), None, None)
    call("riqn_split_bf16", N, K, ptr(b), ptr(b_hi), ptr(b_lo), None, None)
    assert np.array_equal(a_hi.float().cpu().numpy(), _bf16_round(A))
    bias = torch.from_numpy(rs.standard_normal(N).astype(np.float32)).to(dev)
    c = torch.zeros(M, N, device=dev)
    eps = torch.from_numpy(rs.standard_normal((M, N)).astype(np.float32)).to(dev)
    c2 = torch.zeros(M, N, device=dev)
    call("riqn_gemm_bf16_tc", M, N, K, ptr(a_hi), ptr(a_lo), ptr(b_hi), ptr(b_lo), ptr(c), N, epi, ptr(bias), ptr(c2),
         ptr(eps), split_k, None, None)
    torch.cuda.synchronize()
    if split3:
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        tol = 2e-5
    else:
        ref = _bf16_round(A).astype(np.float64) @ _bf16_round(B).astype(np.float64).T
        tol = 1e-5      # fp32 accumulation inside the tensor core over up to 8000 products
    got = c.cpu().numpy()
    if epi == 1:
        ref = np.maximum(ref + bias.cpu().numpy().astype(np.float64), 0)
    assert rel_err(got, ref) < tol, (M, N, K, split3, epi, split_k, rel_err(got, ref))
    if epi == 3:
        assert rel_err(c2.cpu().numpy(), ref * eps.cpu().numpy().astype(np.float64)) < max(tol, 1e-5)


@pytest.mark.parametrize("split3", [False, True])
def test_tc_gemm_single_tile(cuda_dev, split3):
    _run(cuda_dev, 128, 256, 64, split3)
    _run(cuda_dev, 128, 256, 256, split3, seed=1)


@pytest.mark.parametrize("split3", [False, True])
def test_tc_gemm_multi_tile_ragged(cuda_dev, split3):
    _run(cuda_dev, 300, 700, 3136, split3, seed=2)            # partial M, N tiles; 49 k-blocks
    _run(cuda_dev, 1000, 1024, 3136, split3, epi=1, seed=3)   # head forward shape (rows x 1024 x 3136) + bias/relu


@pytest.mark.parametrize("split3", [False, True])
def test_tc_gemm_splitk_atomic(cuda_dev, split3):
    _run(cuda_dev, 1024, 3136, 4096, split3, epi=2, split_k=4, seed=4)     # wgrad shape, K = rows
    _run(cuda_dev, 256, 512, 1000 * 8, split3, epi=3, split_k=7, seed=5)   # uneven split + dsigma output


def test_tc_gemm_many_tiles_persistent(cuda_dev):
    _run(cuda_dev, 4096, 2048, 1024, False, seed=6)           # 256 tiles > 148 SMs: accumulator ring wraps
    _run(cuda_dev, 8192, 1024, 512, True, epi=1, seed=7)


def _run_mn(dev, M, N, K, epi=0, split_k=1, alpha=1.0, seed=0, a_is_km=1):
    """C (+)= alpha * A^T B with A (K, M), B (K, N) row-major bf16 -- the MN-major operand mode; a_is_km = 0:
    C (+)= alpha * A B with A (M, K) row-major (K-major A, MN-major B: a data gradient from the untransposed weight)."""
    from rainbow_iqn_apex_b200._lib import call, ptr
    rs = np.random.RandomState(seed)
    A = _bf16_round(rs.standard_normal((K, M) if a_is_km else (M, K)).astype(np.float32))
    B = _bf16_round((rs.standard_normal((K, N)) * 0.05).astype(np.float32))
    a = torch.from_numpy(A).to(dev).to(torch.bfloat16)
    b = torch.from_numpy(B).to(dev).to(torch.bfloat16)
    c0 = rs.standard_normal((M, N)).astype(np.float32) if epi else np.zeros((M, N), np.float32)
    c = torch.from_numpy(c0).to(dev)
    eps = torch.from_numpy(rs.standard_normal((M, N)).astype(np.float32)).to(dev)
    c2 = torch.zeros(M, N, device=dev)
    call("riqn_gemm_bf16_tc_mn", M, N, K, ptr(a), ptr(b), a_is_km, ptr(c), N, epi, ptr(c2), ptr(eps), alpha, split_k)
    torch.cuda.synchronize()
    prod = (A.astype(np.float64).T if a_is_km else A.astype(np.float64)) @ B.astype(np.float64)
    ref = prod if epi == 0 else c0 + alpha * prod
    assert rel_err(c.cpu().numpy(), ref) < 1e-5, (M, N, K, epi, split_k, rel_err(c.cpu().numpy(), ref))
    if epi == 3:
        assert rel_err(c2.cpu().numpy(), alpha * prod * eps.cpu().numpy().astype(np.float64)) < 1e-5


def test_tc_gemm_mn_major(cuda_dev):
    _run_mn(cuda_dev, 128, 256, 64)                                   # one tile, one k-block
    _run_mn(cuda_dev, 128, 256, 512, seed=1)
    _run_mn(cuda_dev, 64, 64, 200, seed=2)                            # narrow tile, ragged reduction
    _run_mn(cuda_dev, 1024, 3136, 4096, epi=3, split_k=4, seed=3)     # NoisyLinear weight gradient shape
    _run_mn(cuda_dev, 32, 576, 2000, epi=2, split_k=5, alpha=0.5, seed=4)   # conv weight gradient shape (Cout x K)
    # mixed majors: A (M, K) K-major, B (K, N) MN-major -- dX = dY W from the untransposed weight
    _run_mn(cuda_dev, 300, 3136, 1024, seed=5, a_is_km=0)
    _run_mn(cuda_dev, 128, 256, 64, seed=6, a_is_km=0)
