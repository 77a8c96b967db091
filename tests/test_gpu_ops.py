"""GPU: per-op parity of the CUDA kernels (through the C-ABI) against the CPU oracle."""
import numpy as np
import pytest
import torch

from helpers import load_params, make_args, rel_err
from oracle import cases, losses, network as net

pytestmark = pytest.mark.gpu


def _call():
    from rainbow_iqn_apex_b200._lib import call, ptr
    return call, ptr


@pytest.mark.parametrize("M,N,K,ta,tb", [(130, 70, 33, False, False), (257, 129, 64, True, False),
                                         (64, 300, 1000, False, True), (19, 1024, 777, True, True)])
def test_gemm_f32_strided(cuda_dev, M, N, K, ta, tb):
    call, ptr = _call()
    rs = np.random.RandomState(0)
    A = rs.standard_normal((M, K)).astype(np.float32)
    B = rs.standard_normal((N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    a = torch.from_numpy(A.T.copy() if ta else A).to(cuda_dev)
    b = torch.from_numpy(B.T.copy() if tb else B).to(cuda_dev)
    c = torch.empty(M, N, device=cuda_dev)
    sa = (1, M) if ta else (K, 1)
    sb = (1, N) if tb else (K, 1)
    call("riqn_gemm_f32", M, N, K, ptr(a), sa[0], sa[1], ptr(b), sb[0], sb[1], ptr(c), N)
    assert rel_err(c.cpu().numpy(), ref) < 1e-5


@pytest.fixture(scope="module")
def nets(cuda_dev):
    from rainbow_iqn_apex_b200.model import DQN
    params = net.make_params(7)
    d = DQN(make_args(cuda_dev), 18).to(cuda_dev)
    load_params(d, params)
    return d, params


def test_trunk_u8_and_f32(cuda_dev, nets):
    d, params = nets
    b = cases.make_batch(3, 6)
    p = net.to_torch(params)
    ref = net.conv_trunk(p, torch.from_numpy(b["states"]).float().div_(255)).numpy()
    got_u8 = d.trunk(torch.from_numpy(b["states"]).to(cuda_dev)).cpu().numpy()
    got_f32 = d.trunk(torch.from_numpy(b["states"]).float().div_(255).to(cuda_dev)).cpu().numpy()
    assert rel_err(got_u8, ref) < 3e-5      # split-bf16x3 tensor-core convolution
    # uint8 ingest folds the /255 into the weights (pixel values are exact bf16 operands); fp32 ingest splits x/255
    assert rel_err(got_u8, got_f32) < 1e-5
    # strided window view (B, 7, 84, 84)[:, 3:7]
    win = torch.from_numpy(np.concatenate([b["states"][:, :3], b["next_states"]], axis=1)).to(cuda_dev)
    got_view = d.trunk(win[:, 3:7]).cpu().numpy()
    ref2 = net.conv_trunk(p, torch.from_numpy(b["next_states"]).float().div_(255)).numpy()
    assert rel_err(got_view, ref2) < 3e-5


def test_forward_injected(cuda_dev, nets):
    d, params = nets
    B, Nq = 5, 8
    b = cases.make_batch(4, B)
    noise = net.make_noise(11)
    tau = torch.from_numpy(np.random.RandomState(5).uniform(0, 1, (Nq * B, 1)).astype(np.float32))
    p = net.apply_noise(net.to_torch(params), noise)
    keep = {}
    ref = net.dqn_forward_iqn(p, torch.from_numpy(b["states"]).float().div_(255), Nq, tau, keep=keep)
    d.train()
    d.reset_noise(noise)
    k2 = {}
    q, tau_out = d.forward(torch.from_numpy(b["states"]).to(cuda_dev), Nq, tau=tau, keep=k2, fresh_weights=True)
    assert torch.equal(tau_out.cpu(), tau)
    tc = k2["tc"]

    def qmajor(t):   # head-internal rows are sample-major (b*Nq + q); the oracle's are quantile-major (q*B + b)
        return t.reshape(B, Nq, -1).transpose(0, 1).reshape(B * Nq, -1)

    cos_gpu = qmajor(tc["cos_hi"].float() + tc["cos_lo"].float()).cpu().numpy()    # bf16 hi + lo images
    assert rel_err(cos_gpu, keep["cos"].numpy()) < 2e-5
    if tc["f16"]:     # default arithmetic: x leaves as fp16(x) (head forward operand) + bf16(x) (backward operand)
        assert tc["x_hi"].dtype == torch.float16
        assert rel_err(qmajor(tc["x_hi"].float()).cpu().numpy(), keep["x"].numpy()) < 5e-4
        assert rel_err(qmajor(tc["x_lo"].float()).cpu().numpy(), keep["x"].numpy()) < 4e-3
        htol = 3e-4
    else:
        x_gpu = qmajor(tc["x_hi"].float() + tc["x_lo"].float()).cpu().numpy()
        assert rel_err(x_gpu, keep["x"].numpy()) < 3e-5
        htol = 1e-4                                                              # split-bf16x3 tensor-core products
    h_gpu = qmajor(k2["h"])
    assert rel_err(h_gpu[:, :512].cpu().numpy(), keep["h_v"].numpy()) < htol
    assert rel_err(h_gpu[:, 512:].cpu().numpy(), keep["h_a"].numpy()) < htol
    assert rel_err(q.cpu().numpy(), ref.numpy()) < htol
    # stored epsilons == outer product of the injected factors (model.py:39-43), bit for bit
    assert torch.equal(d.fcnoisy_h_a.weight_epsilon.cpu(), torch.outer(noise["fcnoisy_h_a"][1], noise["fcnoisy_h_a"][0]))
    # eval mode uses mu only (model.py:52-53)
    d.eval()
    q_eval, _ = d.forward(torch.from_numpy(b["states"]).to(cuda_dev), Nq, tau=tau)
    ref_eval = net.dqn_forward_iqn(p, torch.from_numpy(b["states"]).float().div_(255), Nq, tau, training=False)
    assert rel_err(q_eval.cpu().numpy(), ref_eval.numpy()) < htol
    d.train()


@pytest.mark.parametrize("B,N,Np,kappa", [(7, 8, 8, 1.0), (33, 64, 64, 1.0), (4, 16, 40, 0.5), (3, 100, 9, 2.0)])
def test_iqn_loss_kernel(cuda_dev, B, N, Np, kappa):
    call, ptr = _call()
    rs = np.random.RandomState(B)
    A = 18
    q_on = torch.from_numpy(rs.standard_normal((N * B, A)).astype(np.float32)).requires_grad_(True)
    q_tg = torch.from_numpy(rs.standard_normal((Np * B, A)).astype(np.float32))
    tau = torch.from_numpy(rs.uniform(0, 1, (N * B, 1)).astype(np.float32))
    actions = torch.from_numpy(rs.randint(0, A, B).astype(np.int64))
    a_star = torch.from_numpy(rs.randint(0, A, B).astype(np.int64))
    returns = torch.from_numpy(rs.standard_normal(B).astype(np.float32))
    nt = torch.from_numpy((rs.uniform(size=B) < 0.8).astype(np.float32))
    g = 0.99 ** 3
    target = (returns[:, None].repeat(Np, 1) + (g * nt[:, None]).repeat(Np, 1)
              * q_tg.gather(1, a_star[:, None].repeat(Np, 1))).reshape(Np, B).t()
    theta = q_on.gather(1, actions[:, None].repeat(N, 1)).reshape(N, B).t()
    ref = losses.iqn_pairwise_loss(theta, target, tau.reshape(N, B).t(), kappa)
    w = torch.from_numpy(rs.uniform(0.1, 1, B).astype(np.float32))
    (w * ref).sum().backward()
    dev = cuda_dev
    loss = torch.empty(B, device=dev)
    dth = torch.empty(N * B, device=dev)
    th_o = torch.empty(B, N, device=dev)
    tg_o = torch.empty(B, Np, device=dev)
    d_in = [t.to(dev) for t in (q_on.detach(), q_tg, tau, actions, a_star, returns, nt)]   # keep alive
    call("riqn_iqn_loss_fwd_bwd", B, N, Np, A, *[ptr(t) for t in d_in], float(g), float(kappa),
         ptr(loss), ptr(dth), ptr(th_o), ptr(tg_o))
    assert np.array_equal(tg_o.cpu().numpy(), target.numpy())       # same fp32 op order as the reference
    assert np.array_equal(th_o.cpu().numpy(), theta.detach().numpy())
    assert rel_err(loss.cpu().numpy(), ref.detach().numpy()) < 1e-5   # loss kernel alone <= 1e-5
    # dtheta[i*B+b] * w[b] == dL/dq_on[i*B+b, actions[b]]
    gref = q_on.grad.gather(1, actions[:, None].repeat(N, 1)).reshape(N, B)
    got = dth.cpu().reshape(N, B) * w[None, :]
    assert rel_err(got.numpy(), gref.numpy()) < 1e-5


def test_argmax_mean(cuda_dev):
    call, ptr = _call()
    rs = np.random.RandomState(1)
    B, K, A = 37, 32, 18
    q = torch.from_numpy(rs.standard_normal((K * B, A)).astype(np.float32))
    ref = q.reshape(K, B, A).mean(0).argmax(1)
    out = torch.empty(B, dtype=torch.int64, device=cuda_dev)
    qd = q.to(cuda_dev)
    call("riqn_argmax_mean", B, K, A, ptr(qd), ptr(out))
    assert torch.equal(out.cpu(), ref)


def test_adam_matches_torch(cuda_dev):
    call, ptr = _call()
    rs = np.random.RandomState(2)
    n = 10007
    p0 = rs.standard_normal(n).astype(np.float32)
    p_ref = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([p_ref], lr=5e-5, eps=3.125e-4)
    p = torch.from_numpy(p0.copy()).to(cuda_dev)
    m = torch.zeros(n, device=cuda_dev)
    v = torch.zeros(n, device=cuda_dev)
    for step in range(1, 4):
        g = (rs.standard_normal(n) * 10.0 ** rs.randint(-6, 1, n)).astype(np.float32)
        p_ref.grad = torch.from_numpy(g.copy())
        opt.step()
        gd = torch.from_numpy(g).to(cuda_dev)
        call("riqn_adam_step", n, ptr(p), ptr(gd), ptr(m), ptr(v), step, 5e-5, 0.9, 0.999, 3.125e-4, 1.0, None)
        assert np.allclose(p.cpu().numpy(), p_ref.detach().numpy(), rtol=0, atol=2.5e-7)   # <= 1 fp32 ulp of |p| < 4
        upd, upd_ref = p.cpu().numpy() - p0, p_ref.detach().numpy() - p0
        assert rel_err(upd, upd_ref) < 5e-3
    assert rel_err(m.cpu().numpy(), opt.state[p_ref]["exp_avg"].numpy()) < 1e-6


def test_device_rng_statistics(cuda_dev):
    call, ptr = _call()
    n = 1 << 20
    u = torch.empty(n, device=cuda_dev)
    call("riqn_fill_uniform", n, 1234, 0, ptr(u), None)
    u2 = torch.empty(n, device=cuda_dev)
    call("riqn_fill_uniform", n, 1234, 1, ptr(u2), None)
    a = u.cpu().numpy().astype(np.float64)
    assert 0 < a.min() and a.max() < 1 and abs(a.mean() - 0.5) < 2e-3 and abs(a.var() - 1 / 12) < 1e-3
    assert abs(np.corrcoef(a, u2.cpu().numpy())[0, 1]) < 5e-3
    z = torch.empty(n, device=cuda_dev)
    call("riqn_noisy_sample", n, 99, 0, ptr(z), None)
    f = z.cpu().numpy().astype(np.float64)
    x = np.sign(f) * f * f                      # invert f(x) = sign(x) sqrt|x|  -> N(0,1)
    assert abs(x.mean()) < 5e-3 and abs(x.var() - 1) < 1e-2 and abs((x ** 4).mean() - 3) < 0.1


def test_noisy_reset_net_matches_per_layer_calls(cuda_dev):
    """riqn_noisy_reset_net (two launches per network) draws and composes exactly what riqn_noisy_sample +
    riqn_noisy_compose produce layer by layer on the same seed / stream ids (model.py:32-43,159-162)."""
    from rainbow_iqn_apex_b200._lib import NoisyLayer
    call, ptr = _call()
    shapes = [(512, 3136), (512, 3136), (1, 512), (18, 512)]
    seed, g = 4242, torch.Generator().manual_seed(5)
    keep, desc = [], (NoisyLayer * len(shapes))()
    for k, (o, i) in enumerate(shapes):
        t = dict(mu=torch.randn(o, i, generator=g), sg=torch.rand(o, i, generator=g), bmu=torch.randn(o, generator=g),
                 bsg=torch.rand(o, generator=g))
        t = {n: v.to(cuda_dev) for n, v in t.items()}
        for n, shp in (("eps", (o, i)), ("beps", (o,)), ("ein", (i,)), ("eout", (o,)), ("w", (o, i)), ("b", (o,))):
            t[n] = torch.zeros(*shp, device=cuda_dev)
            t[n + "_ref"] = torch.zeros(*shp, device=cuda_dev)
        keep.append(t)
        d = desc[k]
        d.out_features, d.in_features = o, i
        d.weight_mu, d.weight_sigma, d.weight_epsilon = ptr(t["mu"]), ptr(t["sg"]), ptr(t["eps"])
        d.bias_mu, d.bias_sigma, d.bias_epsilon = ptr(t["bmu"]), ptr(t["bsg"]), ptr(t["beps"])
        d.eps_in, d.eps_out, d.w_eff, d.b_eff = ptr(t["ein"]), ptr(t["eout"]), ptr(t["w"]), ptr(t["b"])
        d.stream_in, d.stream_out = (k << 40) + 6, (k << 40) + 7
    call("riqn_noisy_reset_net", len(shapes), desc, seed, 1, 1, None)
    for k, (o, i) in enumerate(shapes):
        t = keep[k]
        call("riqn_noisy_sample", i, seed, (k << 40) + 6, ptr(t["ein_ref"]), None)
        call("riqn_noisy_sample", o, seed, (k << 40) + 7, ptr(t["eout_ref"]), None)
        call("riqn_noisy_compose", o, i, ptr(t["mu"]), ptr(t["sg"]), ptr(t["eps_ref"]), ptr(t["ein_ref"]),
             ptr(t["eout_ref"]), ptr(t["bmu"]), ptr(t["bsg"]), ptr(t["beps_ref"]), ptr(t["w_ref"]), ptr(t["b_ref"]), 1)
        for n in ("ein", "eout", "eps", "beps", "w", "b"):
            assert torch.equal(t[n], t[n + "_ref"]), (k, n)
        # and it is the reference formula: W = mu + sigma * (eps_out (x) eps_in)
        assert torch.equal(t["w"], t["mu"] + t["sg"] * torch.outer(t["eout"], t["ein"]))
    # eval mode: the effective weights are the means; sample = 0 keeps the given factor vectors
    ein0 = keep[0]["ein"].clone()
    call("riqn_noisy_reset_net", len(shapes), desc, seed + 1, 0, 0, None)
    assert torch.equal(keep[0]["ein"], ein0) and torch.equal(keep[0]["w"], keep[0]["mu"])
    assert torch.equal(keep[3]["b"], keep[3]["bmu"])


def test_split_bf16_multi_matches_single_calls(cuda_dev):
    """riqn_split_bf16_multi (the per-step refresh of the noise-free weight images in one launch) writes exactly what
    riqn_split_bf16 / riqn_split_bf16_scaled write on the (column-permuted) source."""
    from rainbow_iqn_apex_b200._lib import SplitJob
    call, ptr = _call()
    g = torch.Generator().manual_seed(3)
    bf = lambda *sh: torch.zeros(*sh, dtype=torch.bfloat16, device=cuda_dev)
    specs = []
    for rows, cols, div, with_perm, with_t in ((32, 256, 255.0, True, False), (64, 512, 1.0, True, False),
                                               (64, 576, 1.0, False, True), (3136, 64, 1.0, False, False)):
        src = torch.randn(rows, cols, generator=g).to(cuda_dev)
        perm = torch.randperm(cols, generator=g).to(torch.int32).to(cuda_dev) if with_perm else None
        specs.append((src, perm, div, bf(rows, cols), bf(rows, cols), bf(cols, rows) if with_t else None))
    arr = (SplitJob * len(specs))()
    for j, (src, perm, div, hi, lo, hiT) in zip(arr, specs):
        j.src, j.perm, j.rows, j.cols, j.div = ptr(src), ptr(perm), src.shape[0], src.shape[1], div
        j.hi, j.lo, j.hi_t = ptr(hi), ptr(lo), ptr(hiT)
    call("riqn_split_bf16_multi", len(specs), arr)
    for src, perm, div, hi, lo, hiT in specs:
        sp = src[:, perm.long()].contiguous() if perm is not None else src
        rh, rl = bf(*src.shape), bf(*src.shape)
        if div != 1.0:
            call("riqn_split_bf16_scaled", src.shape[0], src.shape[1], ptr(sp), div, ptr(rh), ptr(rl))
        else:
            call("riqn_split_bf16", src.shape[0], src.shape[1], ptr(sp), ptr(rh), ptr(rl), None, None, 0)
        assert torch.equal(hi, rh) and torch.equal(lo, rl)
        if hiT is not None:
            assert torch.equal(hiT, rh.t().contiguous())


def test_trunk_pair_equals_two_trunks(cuda_dev):
    """Online + target conv trunks over the same frames as ONE stacked batch (three launches, two weight sets selected by
    m-tile) == the two trunks run one after the other, bit for bit."""
    from rainbow_iqn_apex_b200 import DQN
    B = 128
    a, b_ = DQN(make_args(cuda_dev), 18).to(cuda_dev), DQN(make_args(cuda_dev), 18).to(cuda_dev)
    load_params(a, net.make_params(31))
    load_params(b_, net.make_params(32))
    x = torch.from_numpy(np.random.RandomState(9).randint(0, 256, (B, 7, 84, 84)).astype(np.uint8)).to(cuda_dev)[:, 3:7]
    pair = a.trunk_pair(b_, x)
    assert pair is not None
    fa, fb = pair
    assert torch.equal(fa, a.trunk(x)) and torch.equal(fb, b_.trunk(x))
    assert not torch.equal(fa, fb)
    ref = net.conv_trunk(net.to_torch(net.make_params(32)), x.cpu().float().div_(255)).numpy()
    assert rel_err(fb.cpu().numpy(), ref) < 3e-5
    assert a.trunk_pair(b_, x[:5]) is None                      # rows per network must fill whole 128-row tiles


@pytest.mark.parametrize("R,B,A", [(8203, 1, 18), (4096, 64, 6), (16384, 512, 18)])
def test_dueling_fwd_streamed_rows(cuda_dev, R, B, A):
    """riqn_dueling_fwd takes the shared-memory-streamed kernel from 4096 rows up (model.py:153-156).  Same operation
    order per output as the register kernel: bit-identical to that kernel run on < 4096-row slices (ragged tail
    included), and equal to the float64 product within fp32 accumulation error."""
    call, ptr = _call()
    g = torch.Generator(device="cpu").manual_seed(R + A)
    h = torch.randn(R, 1024, generator=g).clamp_(min=0).to(cuda_dev)
    wz = (torch.randn(1 + A, 512, generator=g) * 0.05).to(cuda_dev)
    bz = torch.randn(1 + A, generator=g).to(cuda_dev)
    q = torch.full((R, A), float("nan"), device=cuda_dev)
    call("riqn_dueling_fwd", R, B, 512, A, ptr(h), ptr(wz), ptr(bz), ptr(q))
    h64, w64, b64 = h.double().cpu(), wz.double().cpu(), bz.double().cpu()
    v = h64[:, :512] @ w64[0] + b64[0]
    adv = h64[:, 512:] @ w64[1:].t() + b64[1:]
    ref = (v[:, None] + adv - adv.mean(1, keepdim=True)).view(B, R // B, A).transpose(0, 1).reshape(R, A)   # row = q_idx*B + b
    assert rel_err(q.cpu().numpy(), ref.numpy()) < 5e-6
    if B == 1:      # identity row map: slices of the input are slices of the output
        parts = []
        for lo in range(0, R, 4000):
            n = min(4000, R - lo)
            qs = torch.empty(n, A, device=cuda_dev)
            call("riqn_dueling_fwd", n, 1, 512, A, ptr(h[lo:lo + n]), ptr(wz), ptr(bz), ptr(qs))
            parts.append(qs)
        assert torch.equal(q, torch.cat(parts))
