"""Dense-layer GEMM parity: exact-fp32 FFMA kernel and the tcgen05 TF32 kernel vs a float64 host product,
for the three operand-major modes of the step (forward, dgrad, wgrad) and their epilogues."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B, transA, transB):
    a = A.double().cpu()
    b = B.double().cpu()
    a = a.t() if transA else a
    b = b.t() if transB else b
    return a @ b


def _mk(rows, cols, ld, seed):
    g = torch.Generator().manual_seed(seed)
    buf = torch.randn(rows, ld, generator=g)
    return buf.cuda(), buf[:, :cols]


CASES = [
    # (M, N, K) as the math sees them
    (300, 200, 96), (128, 128, 32), (2080, 768, 256), (65, 293, 128), (2080, 256, 1024), (7, 5, 8), (513, 300, 40),
    (1000, 128, 296), (768, 256, 5000), (4160, 1024, 256),
]


def _impl_fn(impl):
    """f32 = exact FFMA kernel; tc1 = tcgen05 1-CTA 128x128 tiles; tc2 = tcgen05 CTA-pair (cta_group::2) tiles."""
    from musicstyletransfer_b200 import ops
    if impl == "f32":
        return ops.gemm
    ops.gemm_tc_set_pair(impl == "tc2")
    return ops.gemm_tc


@pytest.fixture(autouse=True)
def _restore_pair_mode():
    yield
    from musicstyletransfer_b200 import ops
    ops.gemm_tc_set_pair(True)


@pytest.mark.parametrize("impl", ["f32", "tc1", "tc2"])
@pytest.mark.parametrize("mode", ["fwd", "dgrad", "wgrad"])
@pytest.mark.parametrize("M,N,K", CASES)
def test_gemm_modes(impl, mode, M, N, K):
    from musicstyletransfer_b200 import ops
    pad = lambda n: (n + 3) // 4 * 4 + 4
    if mode == "fwd":      # C = A[M,K] * B[N,K]^T
        transA, transB = 0, 1
        Ad, Av = _mk(M, K, pad(K), 1)
        Bd, Bv = _mk(N, K, pad(K), 2)
    elif mode == "dgrad":  # C = A[M,K] * B[K,N]
        transA, transB = 0, 0
        Ad, Av = _mk(M, K, pad(K), 1)
        Bd, Bv = _mk(K, N, pad(N), 2)
    else:                  # C = A[K,M]^T * B[K,N]
        transA, transB = 1, 0
        Ad, Av = _mk(K, M, pad(M), 1)
        Bd, Bv = _mk(K, N, pad(N), 2)
    ldc = pad(N)
    want = _ref(Av, Bv, transA, transB)
    C = torch.full((M, ldc), 7.0, device="cuda")
    fn = _impl_fn(impl)
    fn(Ad, Ad.shape[1], transA, Bd, Bd.shape[1], transB, C, ldc, M, N, K)
    torch.cuda.synchronize()
    got = C[:, :N].double().cpu()
    scale = float(want.abs().max()) + 1e-9
    err = float((got - want).abs().max()) / scale
    assert err < ((2e-6 if K <= 1024 else 6e-6) if impl == "f32" else 2e-3), (impl, mode, M, N, K, err)
    # padding columns untouched; the TMA-store epilogue of the tensor path writes whole 16-byte chunks, so it may
    # clobber columns N .. roundup4(N)-1 (documented in include/msx.h) but nothing beyond
    first_safe = N if impl == "f32" else (N + 3) // 4 * 4
    assert float((C[:, first_safe:] - 7.0).abs().max()) == 0.0
    # split-K accumulation into a pre-zeroed C
    C2 = torch.zeros((M, ldc), device="cuda")
    fn(Ad, Ad.shape[1], transA, Bd, Bd.shape[1], transB, C2, ldc, M, N, K, splitk=3)
    torch.cuda.synchronize()
    err = float((C2[:, :N].double().cpu() - want).abs().max()) / scale
    assert err < ((4e-6 if K <= 1024 else 8e-6) if impl == "f32" else 2e-3), ("splitk", impl, mode, err)


@pytest.mark.parametrize("impl", ["f32", "tc1", "tc2"])
def test_gemm_epilogues(impl):
    from musicstyletransfer_b200 import ops
    fn = _impl_fn(impl)
    tol = 2e-6 if impl == "f32" else 2e-3
    M, N, K = 333, 293, 128
    ldc = 296
    Ad, Av = _mk(M, K, K, 3)
    Bd, Bv = _mk(N, K, K, 4)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(5)).cuda()
    want = _ref(Av, Bv, 0, 1) + bias.double().cpu()
    # bias + relu
    C = torch.zeros((M, ldc), device="cuda")
    fn(Ad, K, 0, Bd, K, 1, C, ldc, M, N, K, bias=bias, relu=True)
    got = C[:, :N].double().cpu()
    scale = float(want.abs().max())
    assert float((got - want.clamp(min=0)).abs().max()) / scale < tol
    # accumulate
    C0 = torch.randn(M, ldc, generator=torch.Generator().manual_seed(6)).cuda()
    C = C0.clone()
    fn(Ad, K, 0, Bd, K, 1, C, ldc, M, N, K, bias=bias, accumulate=True)
    assert float((C[:, :N].double().cpu() - (want + C0[:, :N].double().cpu())).abs().max()) / scale < tol
    # aux mask (relu' * scale)
    aux = torch.randn(M, ldc, generator=torch.Generator().manual_seed(7)).cuda()
    C = torch.zeros((M, ldc), device="cuda")
    fn(Ad, K, 0, Bd, K, 1, C, ldc, M, N, K, aux=aux, ldaux=ldc, aux_scale=1.25)
    w2 = _ref(Av, Bv, 0, 1) * (aux[:, :N].cpu() > 0).double() * 1.25
    assert float((C[:, :N].double().cpu() - w2).abs().max()) / scale < tol
    # dropout: both kernels draw the same Philox mask for (seed, site, element)
    C = torch.zeros((M, ldc), device="cuda")
    fn(Ad, K, 0, Bd, K, 1, C, ldc, M, N, K, bias=bias, drop_p=0.25, seed=1234, site=3)
    got = C[:, :N].double().cpu()
    kept = got != 0
    frac = float(kept.double().mean())
    assert 0.70 < frac < 0.80
    assert float((got[kept] - (want / 0.75)[kept]).abs().max()) / scale < tol * 2
    Cs = torch.zeros((M, ldc), device="cuda")
    ops.gemm(Ad, K, 0, Bd, K, 1, Cs, ldc, M, N, K, bias=bias, drop_p=0.25, seed=1234, site=3)
    assert bool(((Cs[:, :N] != 0) == (C[:, :N] != 0)).all())


def test_wgrad_colsum():
    from musicstyletransfer_b200 import ops
    Mred, Nout, Kout = 1000, 293, 128
    dY, dYv = _mk(Mred, Nout, 296, 8)
    X, Xv = _mk(Mred, Kout, Kout, 9)
    gw = torch.zeros(Nout, Kout, device="cuda")
    gb = torch.zeros(Nout, device="cuda")
    ops.gemm(dY, 296, 1, X, Kout, 0, gw, Kout, Nout, Kout, Mred, splitk=4, colsum=gb)
    torch.cuda.synchronize()
    assert float((gw.double().cpu() - dYv.double().t() @ Xv.double()).abs().max()) < 1e-3
    assert float((gb.double().cpu() - dYv.double().sum(0)).abs().max()) < 1e-3


@pytest.mark.parametrize("impl", ["tc1", "tc2"])
def test_tc_epilogue_out_colsum(impl):
    """dgrad epilogue of the tensor path accumulates the column sums of the gradient it writes (bias gradient)."""
    from musicstyletransfer_b200 import ops
    _impl_fn(impl)
    M, N, K = 777, 300, 96
    Ad, Av = _mk(M, K, K, 21)
    Bd, Bv = _mk(K, N, 304, 22)
    aux = torch.randn(M, 304, generator=torch.Generator().manual_seed(23)).cuda()
    C = torch.zeros((M, 304), device="cuda")
    cs = torch.zeros(N, device="cuda")
    ops.gemm_tc(Ad, K, 0, Bd, 304, 0, C, 304, M, N, K, aux=aux, ldaux=304, aux_scale=1.25, out_colsum=cs)
    torch.cuda.synchronize()
    want = (_ref(Av, Bv, 0, 0) * (aux[:, :N].cpu() > 0).double() * 1.25).sum(0)
    assert float((cs.double().cpu() - want).abs().max()) / float(want.abs().max()) < 3e-3
    assert float((cs.double().cpu() - C[:, :N].double().cpu().sum(0)).abs().max()) < 1e-2


# ------------------------------------------------------------------------------------------------ bf16 variant
def _mk16(rows, cols, ld, seed):
    """bf16 operand made on the device by msx_cast_f32_bf16 + the exactly representable values the GEMM will see."""
    from musicstyletransfer_b200 import ops
    g = torch.Generator().manual_seed(seed)
    buf = torch.randn(rows, ld, generator=g).cuda()
    b16 = torch.empty((rows, ld), dtype=torch.bfloat16, device="cuda")
    ops.cast_bf16(buf, b16)
    torch.cuda.synchronize()
    assert bool((b16 == buf.to(torch.bfloat16)).all())          # the cast kernel rounds to nearest even
    return b16, b16[:, :cols].float()


@pytest.mark.parametrize("impl", ["tc1", "tc2"])
@pytest.mark.parametrize("c16", [False, True])
@pytest.mark.parametrize("mode", ["fwd", "dgrad", "wgrad"])
@pytest.mark.parametrize("M,N,K", CASES)
def test_gemm_bf16_modes(impl, c16, mode, M, N, K):
    """bf16 operands are exact inputs; products are exact in fp32, so only the accumulation order differs from the
    float64 host product: fp32 C agrees to ~1e-5, a bf16 C to its own rounding (2^-8 relative per element)."""
    from musicstyletransfer_b200 import ops
    ops.gemm_tc_set_pair(impl == "tc2")
    pad = lambda n: (n + 7) // 8 * 8 + 8
    if mode == "fwd":
        transA, transB = 0, 1
        Ad, Av = _mk16(M, K, pad(K), 1)
        Bd, Bv = _mk16(N, K, pad(K), 2)
    elif mode == "dgrad":
        transA, transB = 0, 0
        Ad, Av = _mk16(M, K, pad(K), 1)
        Bd, Bv = _mk16(K, N, pad(N), 2)
    else:
        transA, transB = 1, 0
        Ad, Av = _mk16(K, M, pad(M), 1)
        Bd, Bv = _mk16(K, N, pad(N), 2)
    ldc = pad(N)
    want = _ref(Av, Bv, transA, transB)
    C = torch.full((M, ldc), 7.0, device="cuda", dtype=torch.bfloat16 if c16 else torch.float32)
    ops.gemm_tc_bf16(Ad, Ad.shape[1], transA, Bd, Bd.shape[1], transB, C, ldc, M, N, K)
    torch.cuda.synchronize()
    got = C[:, :N].double().cpu()
    scale = float(want.abs().max()) + 1e-9
    err = float((got - want).abs().max()) / scale
    assert err < (5e-3 if c16 else 2e-5), (impl, c16, mode, M, N, K, err)
    first_safe = (N + 7) // 8 * 8 if c16 else (N + 3) // 4 * 4
    assert float((C[:, first_safe:].float() - 7.0).abs().max()) == 0.0
    if not c16:
        C2 = torch.zeros((M, ldc), device="cuda")
        ops.gemm_tc_bf16(Ad, Ad.shape[1], transA, Bd, Bd.shape[1], transB, C2, ldc, M, N, K, splitk=3)
        torch.cuda.synchronize()
        err = float((C2[:, :N].double().cpu() - want).abs().max()) / scale
        assert err < 2e-5, ("splitk", impl, mode, err)


@pytest.mark.parametrize("impl", ["tc1", "tc2"])
def test_gemm_bf16_epilogues(impl):
    from musicstyletransfer_b200 import ops
    ops.gemm_tc_set_pair(impl == "tc2")
    M, N, K = 333, 293, 256
    ldc = 296
    Ad, Av = _mk16(M, K, K, 3)
    Bd, Bv = _mk16(N, K, K, 4)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(5)).cuda()
    want = _ref(Av, Bv, 0, 1) + bias.double().cpu()
    scale = float(want.abs().max())
    # bias + relu into a bf16 C (the FF hidden activation of the bf16 step)
    C = torch.zeros((M, ldc), device="cuda", dtype=torch.bfloat16)
    ops.gemm_tc_bf16(Ad, K, 0, Bd, K, 1, C, ldc, M, N, K, bias=bias, relu=True)
    assert float((C[:, :N].double().cpu() - want.clamp(min=0)).abs().max()) / scale < 5e-3
    # fp32 accumulate
    C0 = torch.randn(M, ldc, generator=torch.Generator().manual_seed(6)).cuda()
    Cf = C0.clone()
    ops.gemm_tc_bf16(Ad, K, 0, Bd, K, 1, Cf, ldc, M, N, K, bias=bias, accumulate=True)
    assert float((Cf[:, :N].double().cpu() - (want + C0[:, :N].double().cpu())).abs().max()) / scale < 2e-5
    # bf16 aux mask (relu' * scale) + column sums of what is written, bf16 C
    aux32 = torch.randn(M, ldc, generator=torch.Generator().manual_seed(7)).cuda().clamp(min=0)
    aux = aux32.to(torch.bfloat16)
    for cdt in (torch.float32, torch.bfloat16):
        C = torch.zeros((M, ldc), device="cuda", dtype=cdt)
        cs = torch.zeros(N, device="cuda")
        ops.gemm_tc_bf16(Ad, K, 0, Bd, K, 1, C, ldc, M, N, K, aux=aux, ldaux=ldc, aux_scale=1.25, out_colsum=cs)
        w2 = _ref(Av, Bv, 0, 1) * (aux[:, :N].float().cpu() > 0).double() * 1.25
        assert float((C[:, :N].double().cpu() - w2).abs().max()) / scale < (5e-3 if cdt == torch.bfloat16 else 2e-5)
        assert float((cs.double().cpu() - w2.sum(0)).abs().max()) / float(w2.sum(0).abs().max()) < 1e-4
    # fp32 aux with a bf16 GEMM, unaligned tail columns (N = 293 is not a multiple of 32)
    C = torch.zeros((M, ldc), device="cuda")
    ops.gemm_tc_bf16(Ad, K, 0, Bd, K, 1, C, ldc, M, N, K, aux=aux32, ldaux=ldc, aux_scale=2.0)
    w3 = _ref(Av, Bv, 0, 1) * (aux32[:, :N].cpu() > 0).double() * 2.0
    assert float((C[:, :N].double().cpu() - w3).abs().max()) / scale < 2e-5
    # dropout: same counter-hash mask as the fp32-operand kernels for (seed, site, element)
    C = torch.zeros((M, ldc), device="cuda", dtype=torch.bfloat16)
    ops.gemm_tc_bf16(Ad, K, 0, Bd, K, 1, C, ldc, M, N, K, bias=bias, drop_p=0.25, seed=1234, site=3)
    Af, Bf = Ad.float().contiguous(), Bd.float().contiguous()
    Cs = torch.zeros((M, ldc), device="cuda")
    ops.gemm(Af, K, 0, Bf, K, 1, Cs, ldc, M, N, K, bias=bias, drop_p=0.25, seed=1234, site=3)
    torch.cuda.synchronize()
    assert float((C[:, :N].float() - Cs[:, :N]).abs().max()) / scale < 8e-3


@pytest.mark.parametrize("impl", ["tc1", "tc2"])
@pytest.mark.parametrize("ab16", [False, True])
def test_gemm_relu_bit_mask(impl, ab16):
    """msx_gemm_tc_ex: the forward epilogue writes the bit mask of (C > 0) after bias / ReLU / dropout, and a dgrad that
    takes that mask as aux (4 bytes per 32 elements) equals the dgrad that re-reads C itself."""
    from musicstyletransfer_b200 import ops
    ops.gemm_tc_set_pair(impl == "tc2")
    M, N, K = 333, 256, 256
    mk = _mk16 if ab16 else (lambda r, c, ld, s: _mk(r, c, ld, s))
    fn = ops.gemm_tc_bf16 if ab16 else ops.gemm_tc
    Ad, Av = mk(M, K, K, 31)
    Bd, Bv = mk(N, K, K, 32)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(33)).cuda()
    for cdt in ([torch.float32, torch.bfloat16] if ab16 else [torch.float32]):
        C = torch.zeros((M, N), device="cuda", dtype=cdt)
        mask = torch.full((M, N // 32), -1, device="cuda", dtype=torch.int32)
        fn(Ad, K, 0, Bd, K, 1, C, N, M, N, K, bias=bias, relu=True, drop_p=0.3, seed=77, site=1, mask_out=mask, ldmask=N // 32)
        torch.cuda.synchronize()
        bits = ((mask.cpu().numpy().astype(np.uint32)[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(M, N)
        pos = (C.float() > 0).cpu().numpy()
        if cdt == torch.float32:
            assert (bits == pos).all()
        else:       # the mask is taken on the fp32 value before the bf16 rounding: tiny positives may round to +0
            assert (bits >= pos).all() and (bits != pos).mean() < 1e-3
        # dgrad dX = (dY W) * mask * scale: bit mask vs the matrix itself
        dY, _ = mk(M, K, K, 34)            # [M, K'] with K' = K as the reduction dim
        W, _ = mk(K, N, N, 35)             # [K', N] stored as written (transB = 0)
        out_a = torch.zeros((M, N), device="cuda", dtype=cdt)
        out_b = torch.zeros((M, N), device="cuda", dtype=cdt)
        cs_a, cs_b = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
        Cpos = (torch.from_numpy(bits.astype(np.float32)).cuda()).to(cdt).contiguous()      # same mask as a matrix
        fn(dY, K, 0, W, N, 0, out_a, N, M, N, K, aux=Cpos, ldaux=N, aux_scale=1.25, out_colsum=cs_a)
        fn(dY, K, 0, W, N, 0, out_b, N, M, N, K, aux=mask, ldaux=N // 32, aux_scale=1.25, out_colsum=cs_b)
        torch.cuda.synchronize()
        assert bool((out_a == out_b).all())
        assert float((cs_a - cs_b).abs().max()) <= 1e-3 * float(cs_a.abs().max())
