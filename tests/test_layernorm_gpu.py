"""Fused (dropout) + residual + LayerNorm kernels (msx_add_ln_{fwd,bwd}_ex) vs float64 autograd: every vector / scalar
path, the cp.async-prefetching backward (several rows per warp), dropout masks (recovered through the GEMM epilogue,
which draws the same counter-hash mask for the same seed / site / element), bf16 inputs / outputs, accumulation and the
decoder's fused x == y form."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mask(M, D, p, seed, site):
    """keep mask of dropout site `site`: C = dropout(ones[M,4] @ ones[D,4]^T) is non-zero exactly where kept."""
    from musicstyletransfer_b200 import ops
    if p <= 0:
        return torch.ones(M, D, dtype=torch.float64)
    A, Bm = torch.ones(M, 4, device="cuda"), torch.ones(D, 4, device="cuda")
    C = torch.zeros(M, D, device="cuda")
    ops.gemm(A, 4, 0, Bm, 4, 1, C, D, M, D, 4, drop_p=p, seed=seed, site=site)
    torch.cuda.synchronize()
    return (C > 0).double().cpu()


@pytest.mark.parametrize("M,D", [(100, 256), (37, 128), (20, 512), (9, 1024), (50, 32), (50, 64), (8000, 256), (5000, 128)])
@pytest.mark.parametrize("p", [0.0, 0.25])
@pytest.mark.parametrize("y16", [False, True])
def test_add_ln_forward_backward(M, D, p, y16):
    from musicstyletransfer_b200 import ops
    if y16 and D % 128 != 0:
        pytest.skip("bf16 tensors need the vector path (D % 128 == 0)")
    g = torch.Generator().manual_seed(M + D)
    x, y, dout = (torch.randn(M, D, generator=g) for _ in range(3))
    if y16:
        y = y.to(torch.bfloat16).float()            # exactly representable, so the reference sees the same numbers
    gamma, beta = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    seed, site = 1234, 5
    keep = _mask(M, D, p, seed, site) / (1 - p)
    xd, yd, gd = x.double().requires_grad_(True), y.double().requires_grad_(True), gamma.double().requires_grad_(True)
    bd = beta.double().requires_grad_(True)
    s = xd + yd * keep
    mu, var = s.mean(-1, keepdim=True), s.var(-1, unbiased=False, keepdim=True)
    ref = (s - mu) / torch.sqrt(var + 1e-5) * gd + bd
    (ref * dout.double()).sum().backward()

    dev = lambda t, dt=torch.float32: t.to("cuda", dt).contiguous()
    X, Y, G, Bt, DO = dev(x), dev(y, torch.bfloat16 if y16 else torch.float32), dev(gamma), dev(beta), dev(dout)
    out = torch.empty(M, D, device="cuda")
    out16 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16) if D % 128 == 0 else None
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    ops.add_ln_fwd(X, Y, G, Bt, out, mean, rstd, M, D, drop_p=p, seed=seed, site=site, out16=out16)
    torch.cuda.synchronize()
    assert float((out.double().cpu() - ref.detach()).abs().max()) < 2e-5 * float(ref.detach().abs().max())
    if out16 is not None:
        assert bool((out16 == out.to(torch.bfloat16)).all())

    dres, dy = torch.full((M, D), 7.0, device="cuda"), torch.full((M, D), 7.0, device="cuda")
    dy16 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16) if D % 128 == 0 else None
    dgam, dbet, dyb = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    ops.add_ln_bwd(X, Y, G, mean, rstd, DO, dres, dy, dgam, dbet, M, D, drop_p=p, seed=seed, site=site, dybias=dyb, dy16=dy16)
    torch.cuda.synchronize()
    sc = float(xd.grad.abs().max())
    assert float((dres.double().cpu() - xd.grad).abs().max()) < 3e-5 * sc
    assert float((dy.double().cpu() - yd.grad).abs().max()) < 3e-5 * max(sc, float(yd.grad.abs().max()))
    if dy16 is not None:
        assert bool((dy16 == dy.to(torch.bfloat16)).all())
    tol = 2e-4 if M > 1000 else 3e-5                         # long fp32 column sums
    assert float((dgam.double().cpu() - gd.grad).abs().max()) < tol * float(gd.grad.abs().max())
    assert float((dbet.double().cpu() - bd.grad).abs().max()) < tol * float(bd.grad.abs().max())
    assert float((dyb.double().cpu() - yd.grad.sum(0)).abs().max()) < tol * float(yd.grad.sum(0).abs().max()) + 1e-6
    # accumulate_dres adds onto the existing residual gradient (direct-load kernel)
    dres2 = torch.full((M, D), 2.0, device="cuda")
    dg2, db2 = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    ops.add_ln_bwd(X, Y, G, mean, rstd, DO, dres2, None, dg2, db2, M, D, drop_p=p, seed=seed, site=site, accumulate_dres=True)
    torch.cuda.synchronize()
    assert float((dres2.double().cpu() - 2.0 - xd.grad).abs().max()) < 3e-5 * sc + 1e-6


@pytest.mark.parametrize("M,D,p", [(64, 128, 0.0), (300, 256, 0.3)])
def test_add_ln_fused_xy(M, D, p):
    """The decoder's ln3(f + drop(f)): x and y alias, one combined gradient ds * (1 + keep)."""
    from musicstyletransfer_b200 import ops
    g = torch.Generator().manual_seed(D)
    f, dout = torch.randn(M, D, generator=g), torch.randn(M, D, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    seed, site = 99, 130
    keep = _mask(M, D, p, seed, site) / (1 - p)
    fd = f.double().requires_grad_(True)
    s = fd + fd * keep
    mu, var = s.mean(-1, keepdim=True), s.var(-1, unbiased=False, keepdim=True)
    ref = (s - mu) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()
    (ref * dout.double()).sum().backward()
    F, G, Bt, DO = f.cuda(), gamma.cuda(), beta.cuda(), dout.cuda()
    out, mean, rstd = torch.empty(M, D, device="cuda"), torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    ops.add_ln_fwd(F, F, G, Bt, out, mean, rstd, M, D, drop_p=p, seed=seed, site=site)
    df = torch.empty(M, D, device="cuda")
    df16 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    dg, db, dyb = (torch.zeros(D, device="cuda") for _ in range(3))
    ops.add_ln_bwd(F, F, G, mean, rstd, DO, df, None, dg, db, M, D, drop_p=p, seed=seed, site=site, fuse_xy=True, dybias=dyb,
                   dy16=df16)
    torch.cuda.synchronize()
    assert float((out.double().cpu() - ref.detach()).abs().max()) < 2e-5 * float(ref.detach().abs().max())
    sc = float(fd.grad.abs().max())
    assert float((df.double().cpu() - fd.grad).abs().max()) < 3e-5 * sc
    assert bool((df16 == df.to(torch.bfloat16)).all())
    assert float((dyb.double().cpu() - fd.grad.sum(0)).abs().max()) < 1e-4 * float(fd.grad.sum(0).abs().max()) + 1e-6
