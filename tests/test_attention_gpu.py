"""Attention kernels (reference convention: softmax over the query axis) vs a float64 host computation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_fwd(qkv, mask, B, T, H, dh):
    D = H * dh
    x = qkv.double().view(B, T, 3, H, dh)
    K, Q, V = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)   # [B,H,T,dh]
    S = K @ Q.transpose(-1, -2) / np.sqrt(dh)
    m = torch.where(mask.view(B, T) > 0, 0.0, -1e9).double()
    # fp32 semantics of the reference: S + (-1e9) rounds to exactly -1e9 (uniform softmax over the row) while
    # autograd still passes the gradient through the add -> straight-through value replacement
    pad = (m[:, None, :, None] < 0).to(S.dtype)
    S = S + pad * (-1e9 - S).detach()
    P = torch.softmax(S, dim=-1)
    O = P.transpose(-1, -2) @ V
    return O.transpose(1, 2).reshape(B * T, D)


def _inputs(B, T, H, dh, seed):
    g = torch.Generator().manual_seed(seed)
    qkv = torch.randn(B * T, 3 * H * dh, generator=g)
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[0] = T
    mask = (torch.arange(T)[None, :] < lens[:, None]).float().reshape(-1)
    return qkv, mask


# the last two cases give every persistent pipeline group several (batch, head) items (mbarrier phases wrap)
@pytest.mark.parametrize("B,T,H", [(3, 65, 8), (5, 66, 2), (2, 16, 4), (4, 97, 3), (2, 128, 2), (7, 1, 2), (3, 2, 8),
                                   (70, 65, 8), (300, 65, 8), (500, 97, 2), (400, 120, 2)])
def test_attention_tc_forward(B, T, H):
    from musicstyletransfer_b200 import ops
    dh = 32
    qkv, mask = _inputs(B, T, H, dh, seed=T)
    want = _ref_fwd(qkv, mask, B, T, H, dh)
    qd, md = qkv.cuda(), mask.cuda()
    assert ops.attention_tc_supported(qd, T, dh)
    ctx = torch.full((B * T, H * dh), 3.0, device="cuda")
    ops.attention_tc_fwd(qd, md, ctx, B, T, H, dh)
    ctx2 = torch.empty((B * T, H * dh), device="cuda")
    ops.attention_fwd(qd, md, ctx2, B, T, H, dh)
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    assert float((ctx2.double().cpu() - want).abs().max()) / scale < 1e-5
    err = float((ctx.double().cpu() - want).abs().max()) / scale
    assert err < 3e-3, err


@pytest.mark.parametrize("B,T,H,dh", [(3, 65, 8, 32), (300, 65, 8, 32), (5, 66, 2, 32), (2, 16, 4, 32), (4, 97, 3, 32),
                                      (2, 128, 2, 32), (7, 1, 2, 32), (500, 97, 2, 32), (64, 66, 8, 16)])
def test_attention_tc_forward_compensated_scores(B, T, H, dh):
    """x3_scores: S = K Q^T with 3xTF32 operand splitting.  Scores are scaled up (|S| ~ 10) so that the softmax turns the
    single-pass TF32 score error into a visible error of the output; the compensated kernel must be >= 10x closer."""
    from musicstyletransfer_b200 import ops
    qkv, mask = _inputs(B, T, H, dh, seed=T + 7)
    qkv[:, :2 * H * dh] *= 1.8                      # K and Q: score standard deviation ~ 3.2 * sqrt(dh) / sqrt(dh)
    want = _ref_fwd(qkv, mask, B, T, H, dh)
    qd, md = qkv.cuda(), mask.cuda()
    outs = []
    for x3 in (False, True):
        ctx = torch.full((B * T, H * dh), 3.0, device="cuda")
        ops.attention_tc_fwd(qd, md, ctx, B, T, H, dh, x3_scores=x3)
        torch.cuda.synchronize()
        outs.append(float((ctx.double().cpu() - want).abs().max()) / float(want.abs().max()))
    print("attention forward B=%d T=%d H=%d dh=%d: TF32 scores err %.2e, compensated %.2e" % (B, T, H, dh, outs[0], outs[1]))
    assert outs[1] < 6e-4, outs                     # what is left: P and V rounded to TF32 in O = P^T V
    assert outs[1] < 0.7 * outs[0] or T == 1, outs   # T == 1: one query, P == 1 whatever the score


@pytest.mark.parametrize("B,T,H", [(3, 65, 8), (300, 65, 8), (5, 66, 2), (2, 16, 4), (4, 97, 3), (2, 128, 2), (7, 1, 2), (500, 97, 2)])
@pytest.mark.parametrize("x3,planes", [(False, False), (True, False), (True, True)])
def test_attention_tc_forward_query0_only(B, T, H, x3, planes):
    """q0_only (encoder top layer): the context row of query 0 equals the full kernel's / the float64 reference's, every
    other row of the output buffer is left untouched; fp32 output and bf16 hi / lo planes."""
    from musicstyletransfer_b200 import ops
    dh = 32
    qkv, mask = _inputs(B, T, H, dh, seed=T + 3)
    want = _ref_fwd(qkv, mask, B, T, H, dh).view(B, T, H * dh)[:, 0]
    qd, md = qkv.cuda(), mask.cuda()
    if planes:
        hi = torch.full((B * T, H * dh), 3.0, device="cuda", dtype=torch.bfloat16)
        lo = torch.full((B * T, H * dh), 3.0, device="cuda", dtype=torch.bfloat16)
        ops.attention_tc_fwd(qd, md, hi, B, T, H, dh, x3_scores=x3, ctx_lo=lo, q0_only=True)
        torch.cuda.synchronize()
        ctx = (hi.float() + lo.float()).view(B, T, H * dh)
        untouched = 6.0
    else:
        ctx = torch.full((B * T, H * dh), 3.0, device="cuda")
        ops.attention_tc_fwd(qd, md, ctx, B, T, H, dh, x3_scores=x3, q0_only=True)
        torch.cuda.synchronize()
        ctx = ctx.view(B, T, H * dh)
        untouched = 3.0
    err = float((ctx[:, 0].double().cpu() - want).abs().max()) / float(want.abs().max())
    assert err < (6e-4 if x3 else 3e-3), err
    if T > 1:
        assert bool((ctx[:, 1:] == untouched).all())


# T <= 80 runs the pipelined kernel (the larger B cases give every pipeline group several items; H = 3 makes the head
# change between the items of a group, which exercises the bias-gradient flush), T > 80 the one-shot kernel
@pytest.mark.parametrize("B,T,H", [(3, 65, 8), (5, 66, 2), (2, 16, 4), (4, 97, 3), (2, 128, 2), (7, 1, 2), (3, 2, 8),
                                   (70, 65, 8), (300, 65, 8), (220, 66, 3), (400, 33, 2)])
def test_attention_backward(B, T, H):
    """dqkv of both backward kernels vs torch autograd (float64) through the reference formula."""
    from musicstyletransfer_b200 import ops
    dh = 32
    qkv, mask = _inputs(B, T, H, dh, seed=100 + T)
    g = torch.Generator().manual_seed(T)
    dctx = torch.randn(B * T, H * dh, generator=g)
    x = qkv.double().requires_grad_(True)
    (_ref_fwd(x, mask, B, T, H, dh) * dctx.double()).sum().backward()
    want = x.grad
    scale = float(want.abs().max())
    qd, md, dd = qkv.cuda(), mask.cuda(), dctx.cuda()
    out = torch.empty_like(qd)
    ops.attention_bwd(qd, md, dd, out, B, T, H, dh)
    torch.cuda.synchronize()
    assert float((out.double().cpu() - want).abs().max()) / scale < 1e-4
    out2 = torch.full_like(qd, 5.0)
    db = torch.zeros(3 * H * dh, device="cuda")
    ops.attention_tc_bwd(qd, md, dd, out2, B, T, H, dh, dbias=db)
    torch.cuda.synchronize()
    err = float((out2.double().cpu() - want).abs().max()) / scale
    assert err < 5e-3, err
    wb = want.sum(0)
    assert float((db.double().cpu() - wb).abs().max()) <= 5e-3 * float(wb.abs().max()) + 1e-4


@pytest.mark.parametrize("B,T,H", [(3, 65, 8), (5, 66, 2), (2, 16, 4), (4, 97, 3), (7, 1, 2), (3, 2, 8), (70, 65, 8), (300, 65, 8),
                                   (220, 66, 3), (400, 33, 2)])
@pytest.mark.parametrize("out16", [False, True])
def test_attention_backward_query0_only(B, T, H, out16):
    """q0_only: the context gradient is non-zero in the row of query 0 of every sequence only (the encoder's top layer).
    dqkv and the bias gradient vs torch autograd (float64); T > 80 takes the general kernel through the same entry point."""
    from musicstyletransfer_b200 import ops
    dh = 32
    qkv, mask = _inputs(B, T, H, dh, seed=200 + T)
    g = torch.Generator().manual_seed(T + 1)
    dctx = torch.zeros(B, T, H * dh)
    dctx[:, 0] = torch.randn(B, H * dh, generator=g)
    dctx = dctx.view(B * T, H * dh)
    x = qkv.double().requires_grad_(True)
    (_ref_fwd(x, mask, B, T, H, dh) * dctx.double()).sum().backward()
    want = x.grad
    scale = float(want.abs().max())
    qd, md, dd = qkv.cuda(), mask.cuda(), dctx.cuda()
    out = torch.full(qd.shape, 5.0, device="cuda", dtype=torch.bfloat16 if out16 else torch.float32)
    db = torch.zeros(3 * H * dh, device="cuda")
    ops.attention_tc_bwd(qd, md, dd, out, B, T, H, dh, dbias=db, q0_only=True)
    torch.cuda.synchronize()
    err = float((out.double().cpu() - want).abs().max()) / scale
    assert err < (1.2e-2 if out16 else 5e-3), err
    wb = want.sum(0)
    assert float((db.double().cpu() - wb).abs().max()) <= 5e-3 * float(wb.abs().max()) + 1e-4


# ------------------------------------------------------------------------------------------------ long rows (T > 128)
@pytest.mark.parametrize("B,T,H", [(2, 129, 2), (3, 257, 2), (2, 200, 3), (5, 256, 1), (2, 384, 2), (40, 129, 8), (3, 144, 2),
                                   (3, 130, 2), (4, 131, 3), (3, 132, 2), (2, 133, 2), (4, 258, 2), (3, 260, 1), (300, 129, 8),
                                   (2, 385, 2), (3, 388, 1), (2, 400, 2), (2, 512, 1), (3, 513, 2), (2, 516, 1), (2, 600, 2),
                                   (2, 641, 1), (2, 700, 2), (2, 768, 1), (40, 513, 8)])
@pytest.mark.parametrize("out16", [False, True])
def test_attention_long_rows(B, T, H, out16):
    """msx_attention_tcl_fwd / _bwd (key tiles + query chunks of 128; T > 384: the two-sweep forward) vs the float64 reference
    formula and autograd."""
    from musicstyletransfer_b200 import ops
    dh = 32
    qkv, mask = _inputs(B, T, H, dh, seed=T)
    g = torch.Generator().manual_seed(T)
    dctx = torch.randn(B * T, H * dh, generator=g)
    x = qkv.double().requires_grad_(True)
    want = _ref_fwd(x, mask, B, T, H, dh)
    (want * dctx.double()).sum().backward()
    wg = x.grad
    want = want.detach()
    qd, md, dd = qkv.cuda(), mask.cuda(), dctx.cuda()
    assert ops.attention_tcl_supported(qd, T, dh) and not ops.attention_tc_supported(qd, T, dh)
    odt = torch.bfloat16 if out16 else torch.float32
    ctx = torch.full((B * T, H * dh), 3.0, device="cuda", dtype=odt)
    stats = torch.zeros((B * H * T, 2), device="cuda")
    ops.attention_tcl_fwd(qd, md, ctx, stats, B, T, H, dh)
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    err = float((ctx.double().cpu() - want).abs().max()) / scale
    assert err < (8e-3 if out16 else 3e-3), err
    out = torch.full((B * T, 3 * H * dh), 5.0, device="cuda", dtype=odt)
    db = torch.zeros(3 * H * dh, device="cuda")
    ops.attention_tcl_bwd(qd, md, dd, stats, out, B, T, H, dh, dbias=db)
    torch.cuda.synchronize()
    gs = float(wg.abs().max())
    err = float((out.double().cpu() - wg).abs().max()) / gs
    assert err < (1e-2 if out16 else 5e-3), err
    wb = wg.sum(0)
    assert float((db.double().cpu() - wb).abs().max()) <= 5e-3 * float(wb.abs().max()) + 1e-4


@pytest.mark.parametrize("B,T,H", [(3, 129, 2), (2, 130, 3), (3, 132, 2), (2, 133, 2), (2, 200, 2), (3, 256, 1), (3, 257, 2), (2, 260, 2),
                                   (2, 384, 2), (200, 129, 8), (2, 385, 2), (2, 400, 1), (3, 513, 2), (2, 516, 1), (2, 640, 2),
                                   (2, 645, 1), (2, 768, 2)])
@pytest.mark.parametrize("out16", [False, True])
def test_attention_long_rows_q0(B, T, H, out16):
    """q0_only variants of the long-row kernels (the encoder's top layer under SOS-rows-only): the forward writes the context
    row of query 0 of every sequence only, the backward takes a context gradient that is zero outside those rows; both vs
    the float64 reference formula / autograd, and the statistics the forward saves vs the full forward's."""
    from musicstyletransfer_b200 import ops
    dh = 32
    D = H * dh
    qkv, mask = _inputs(B, T, H, dh, seed=T + 7)
    g = torch.Generator().manual_seed(T)
    dctx = torch.zeros(B * T, D)
    dctx[::T] = torch.randn(B, D, generator=g)
    x = qkv.double().requires_grad_(True)
    want = _ref_fwd(x, mask, B, T, H, dh)
    (want * dctx.double()).sum().backward()
    wg = x.grad
    want = want.detach()
    qd, md, dd = qkv.cuda(), mask.cuda(), dctx.cuda()
    odt = torch.bfloat16 if out16 else torch.float32
    ctx = torch.full((B * T, D), 3.0, device="cuda", dtype=odt)
    stats = torch.zeros((B * H * T, 2), device="cuda")
    ops.attention_tcl_fwd(qd, md, ctx, stats, B, T, H, dh, q0_only=True)
    torch.cuda.synchronize()
    got = ctx.double().cpu()
    scale = float(want.abs().max())
    err = float((got[::T] - want[::T]).abs().max()) / scale
    assert err < (8e-3 if out16 else 3e-3), err
    rest = torch.ones(B * T, dtype=torch.bool)
    rest[::T] = False
    assert bool((got[rest] == 3.0).all()), "q0_only forward must leave the other context rows untouched"
    stats_full = torch.zeros_like(stats)
    ops.attention_tcl_fwd(qd, md, torch.empty_like(ctx), stats_full, B, T, H, dh)
    torch.cuda.synchronize()
    # the same row maxima bit for bit; the sums may be taken in another order (T = 261 ... 384: q0_only runs the chunked kernel)
    assert torch.equal(stats[:, 0], stats_full[:, 0])
    assert torch.allclose(stats[:, 1], stats_full[:, 1], rtol=2e-6, atol=0.0)
    out = torch.full((B * T, 3 * D), 5.0, device="cuda", dtype=odt)
    db = torch.zeros(3 * D, device="cuda")
    ops.attention_tcl_bwd(qd, md, dd, stats, out, B, T, H, dh, dbias=db, q0_only=True)
    torch.cuda.synchronize()
    gs = float(wg.abs().max())
    err = float((out.double().cpu() - wg).abs().max()) / gs
    assert err < (1e-2 if out16 else 5e-3), err
    wb = wg.sum(0)
    assert float((db.double().cpu() - wb).abs().max()) <= 5e-3 * float(wb.abs().max()) + 1e-4


@pytest.mark.parametrize("T", [129, 257, 400, 513])
@pytest.mark.parametrize("q0", [False, True])
def test_attention_long_rows_batch_invariance(T, q0):
    """Size-independent property at a bench-like grid (B * H = 4736 CTAs, 32 waves): a (batch, head) item is computed by one
    CTA from its own rows only, so the context / statistics / dqkv of a large batch equal, bit for bit, those of the same rows
    processed in four smaller calls (a race between concurrently resident CTAs or a stale shared-memory tile would show here)."""
    from musicstyletransfer_b200 import ops
    B, H, dh = 592, 8, 32
    D = H * dh
    g = torch.Generator().manual_seed(T)
    qkv = torch.randn(B * T, 3 * D, generator=g).cuda()
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    mask = (torch.arange(T)[None, :] < lens[:, None]).float().reshape(-1).cuda()
    dctx = torch.randn(B * T, D, generator=g)
    if q0:
        keep = torch.zeros(B * T, 1)
        keep[::T] = 1.0
        dctx = dctx * keep
    dctx = dctx.cuda()

    def run(lo, hi):
        n = hi - lo
        q, m, d = qkv[lo * T:hi * T], mask[lo * T:hi * T], dctx[lo * T:hi * T]
        ctx = torch.zeros(n * T, D, device="cuda")
        stats = torch.zeros(n * H * T, 2, device="cuda")
        out = torch.zeros(n * T, 3 * D, device="cuda")
        ops.attention_tcl_fwd(q, m, ctx, stats, n, T, H, dh, q0_only=q0)
        ops.attention_tcl_bwd(q, m, d, stats, out, n, T, H, dh, q0_only=q0)
        return ctx, stats, out

    full = run(0, B)
    parts = [run(i * 148, (i + 1) * 148) for i in range(4)]
    torch.cuda.synchronize()
    for k, name in enumerate(("ctx", "stats", "dqkv")):
        assert torch.equal(full[k], torch.cat([p[k] for p in parts])), name
    assert bool(torch.isfinite(full[2]).all())


@pytest.mark.parametrize("B,T,H", [(3, 129, 2), (2, 200, 2), (3, 257, 2), (2, 384, 1), (2, 513, 2)])
@pytest.mark.parametrize("q0", [False, True])
def test_attention_long_rows_planes(B, T, H, q0):
    """ctx as bf16 hi / lo planes straight from the long-row forward (msx_attention_tcl_fwd_p): hi = rn_bf16(o),
    lo = rn_bf16(o - hi) of the very values the fp32 output holds (16 mantissa bits together)."""
    from musicstyletransfer_b200 import ops
    dh = 32
    D = H * dh
    qkv, mask = _inputs(B, T, H, dh, seed=T + 3)
    qd, md = qkv.cuda(), mask.cuda()
    ctx = torch.zeros(B * T, D, device="cuda")
    stats = torch.zeros(B * H * T, 2, device="cuda")
    ops.attention_tcl_fwd(qd, md, ctx, stats, B, T, H, dh, q0_only=q0)
    planes = torch.full((2, B * T, D), 3.0, device="cuda", dtype=torch.bfloat16)
    stats2 = torch.zeros_like(stats)
    ops.attention_tcl_fwd(qd, md, planes[0], stats2, B, T, H, dh, q0_only=q0, ctx_lo=planes[1])
    torch.cuda.synchronize()
    assert torch.equal(stats, stats2)
    rows = slice(0, None, T) if q0 else slice(None)
    want = ctx[rows]
    hi = want.to(torch.bfloat16)
    lo = (want - hi.float()).to(torch.bfloat16)
    assert torch.equal(planes[0][rows], hi) and torch.equal(planes[1][rows], lo)
    if q0:
        rest = torch.ones(B * T, dtype=torch.bool, device="cuda")
        rest[::T] = False
        assert bool((planes[:, rest] == 3.0).all())


# ------------------------------------------------------------------------------------------------ 16-wide heads
@pytest.mark.parametrize("B,T,H", [(3, 66, 8), (5, 65, 3), (2, 16, 4), (4, 97, 2), (2, 128, 8), (300, 66, 8)])
@pytest.mark.parametrize("out16", [False, True])
def test_attention_tc_half_heads(B, T, H, out16):
    """d_h = 16 (the 128-wide Transformer decoder with 8 heads): the tcgen05 kernels load 32-wide tiles, reduce over and
    store the first 16 columns only; forward, backward and the fused K|Q|V bias gradient vs float64 / autograd."""
    from musicstyletransfer_b200 import ops
    dh = 16
    qkv, mask = _inputs(B, T, H, dh, seed=T + 1)
    g = torch.Generator().manual_seed(T)
    dctx = torch.randn(B * T, H * dh, generator=g)
    x = qkv.double().requires_grad_(True)
    want = _ref_fwd(x, mask, B, T, H, dh)
    (want * dctx.double()).sum().backward()
    wg = x.grad
    want = want.detach()
    qd, md, dd = qkv.cuda(), mask.cuda(), dctx.cuda()
    assert ops.attention_tc_supported(qd, T, dh)
    odt = torch.bfloat16 if out16 else torch.float32
    ctx = torch.full((B * T, H * dh), 3.0, device="cuda", dtype=odt)
    ops.attention_tc_fwd(qd, md, ctx, B, T, H, dh)
    torch.cuda.synchronize()
    err = float((ctx.double().cpu() - want).abs().max()) / float(want.abs().max())
    assert err < (8e-3 if out16 else 3e-3), err
    out = torch.full((B * T, 3 * H * dh), 5.0, device="cuda", dtype=odt)
    db = torch.zeros(3 * H * dh, device="cuda")
    ops.attention_tc_bwd(qd, md, dd, out, B, T, H, dh, dbias=db)
    torch.cuda.synchronize()
    err = float((out.double().cpu() - wg).abs().max()) / float(wg.abs().max())
    assert err < (1e-2 if out16 else 5e-3), err
    wb = wg.sum(0)
    assert float((db.double().cpu() - wb).abs().max()) <= 5e-3 * float(wb.abs().max()) + 1e-4


@pytest.mark.parametrize("B,T,H,dh", [(2, 400, 2, 32), (3, 513, 2, 32), (2, 65, 3, 32), (5, 33, 2, 16), (2, 1, 2, 32), (1, 700, 1, 64),
                                      (2, 200, 4, 8)])
def test_attention_tiled_any_length(B, T, H, dh):
    """Key-tiled exact-fp32 attention (attention_tiled.cu): forward and dqkv vs float64 / autograd for rows beyond every other
    kernel's limit (T = 400, 513, 700) and, for cross-checking, lengths the one-CTA kernel takes as well."""
    from musicstyletransfer_b200 import ops
    qkv, mask = _inputs(B, T, H, dh, seed=T + dh)
    g = torch.Generator().manual_seed(T)
    dctx = torch.randn(B * T, H * dh, generator=g)
    x = qkv.double().requires_grad_(True)
    want = _ref_fwd(x, mask, B, T, H, dh)
    (want * dctx.double()).sum().backward()
    qd, md, dd = qkv.cuda(), mask.cuda(), dctx.cuda()
    ctx = torch.full((B * T, H * dh), 7.0, device="cuda")
    ops.attention_tiled_fwd(qd, md, ctx, B, T, H, dh)
    out = torch.full_like(qd, 7.0)
    ops.attention_tiled_bwd(qd, md, dd, out, B, T, H, dh)
    torch.cuda.synchronize()
    ef = float((ctx.double().cpu() - want.detach()).abs().max()) / float(want.abs().max())
    eb = float((out.double().cpu() - x.grad).abs().max()) / float(x.grad.abs().max())
    assert ef < 1e-5 and eb < 1e-4, (ef, eb)
    if T > 384:                      # the generic entry points route long rows to the tiled kernels
        ctx2 = torch.empty_like(ctx)
        ops.attention_fwd(qd, md, ctx2, B, T, H, dh)
        torch.cuda.synchronize()
        assert float((ctx2 - ctx).abs().max()) <= 1e-5 * float(ctx.abs().max())
