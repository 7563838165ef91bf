"""Parity of the path bench.py actually times: the tensor-core step at the bench shape (B = 2048 / 512 rows, L = 64,
raw Xavier weights AND conditioned sigma, several seeds), the step WITH dropout 0.2 (the device's own masks and eps fed
to the oracle), CUDA-graph replay with dropout, the tensor GEMM at the bench's M = 133 120 rows for the six step shapes,
and the tensor-core LSTM recurrence directly against the oracle's lstm_layer.

Tolerances (BASELINE.json north_star): loss, KL, latent means within 1e-3 relative (max |err| / max |value|, the
convention of tests/test_engine_gpu.py).  Measured on B200 at the bench shape, worst of B = 2048 / 3 x B = 512, raw weights:
    tf32x3f (bench headline: 3xTF32 forward GEMMs + compensated attention scores)   means 1.0e-4, loss 8e-7   -> asserted 3e-4
    fp32x3  (strict fp32: every GEMM 3xTF32, exact attention / LSTM)                 means 9.6e-6, loss 5e-7  -> asserted 1e-4
    tf32    (every product single-pass TF32)                                         means 1.30e-3             -> NOT within
            1e-3: it is reported as a faster variant with this deviation stated, asserted < 2.5e-3
Gradients: 1e-3 of each tensor's scale for fp32 / fp32x3, 5 % (mean 2 %) where the backward GEMMs are single-pass TF32."""
import numpy as np
import pytest
import torch

from oracle import model as om

pytestmark = pytest.mark.gpu

BASE_SEED = 0x5EED0000       # engine.VAEEngine.base_seed


def _dev(t, dtype=torch.int32):
    return t.to(dtype).to("cuda:0").contiguous()


def _rel(a, b):
    a = a.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max())


def _bench_rows(B, L, seed):
    """The generator bench.py uses (synth.token_rows_4_4) + a fixed eps."""
    from musicstyletransfer_b200 import synth
    tok, lens, cls, lab = synth.token_rows_4_4(B, L, seed=seed)
    f = lambda a: torch.from_numpy(a).float()
    eps = torch.randn(B, 256, generator=torch.Generator().manual_seed(seed))
    return f(tok), f(lens), f(cls), f(lab), eps


def _condition_sigma(cfg_o, params):
    Z = cfg_o.latent
    params = {k: v.clone() for k, v in params.items()}
    params["encoder.latent_proj.weight"][Z:] *= 0.05
    params["encoder.latent_proj.bias"][Z:] = 3.0
    return params


def _engine(precision, params, dropout=0.0, dec_type="lstm", seed=0):
    from musicstyletransfer_b200.engine import VAEConfig, VAEEngine
    eng = VAEEngine(VAEConfig(dec_type=dec_type, enc_dropout=dropout, dec_dropout=dropout), "cuda:0", seed=seed,
                    precision=precision)
    eng.arena.load_state(params)
    return eng


FWD_TOL = {"tf32x3f": 3e-4, "bf16x3f": 3e-4, "bf16p3f": 3e-4, "fp32x3": 1e-4, "tf32": 2.5e-3}


@pytest.mark.parametrize("precision", ["tf32x3f", "bf16x3f", "bf16p3f", "fp32x3", "tf32"])
@pytest.mark.parametrize("B,seed,conditioned", [(2048, 0, False), (512, 1, False), (512, 2, False), (512, 3, False),
                                                (512, 1, True), (512, 2, True), (512, 3, True)])
def test_bench_shape_forward_vs_oracle(precision, B, seed, conditioned):
    """Loss / KL / latent means of the bench step shape within 1e-3 of the fp32 oracle, on the raw Xavier weights
    bench.py trains from and with sigma conditioned away from zero."""
    cfg_o = om.Cfg(dec_type="lstm")
    p = om.init_params(cfg_o, seed=seed)
    if conditioned:
        p = _condition_sigma(cfg_o, p)
    tokens, lens, classes, labels, eps = _bench_rows(B, 64, seed=10 + seed)
    eng = _engine(precision, p)
    out = eng.forward(_dev(tokens), _dev(lens), _dev(classes), _dev(labels), eps=_dev(eps, torch.float32))
    torch.cuda.synchronize()
    with torch.no_grad():
        _, ce, kl, _, means, stds = om.step_losses(cfg_o, p, tokens, lens, classes, labels, eps)
    dev = {"ce": _rel(out["ce"], ce), "kl": _rel(out["kl"], kl), "means": _rel(out["means"], means),
           "stds": _rel(out["stds"], stds)}
    print("%s B=%d seed=%d conditioned=%s forward deviation:" % (precision, B, seed, conditioned), dev)
    tol = FWD_TOL[precision]
    assert dev["ce"] < tol and dev["means"] < tol and dev["stds"] < tol, dev
    # KL holds log(sigma^2): on raw weights |sigma| reaches ~1e-4 somewhere in a 2048 x 256 batch, where the KL of that
    # ROW is a rounding-level quantity in any fp32 implementation; compare the batch total there and per row otherwise
    if conditioned:
        assert dev["kl"] < tol, dev
    else:
        tot = abs(float(out["kl"].double().sum().cpu()) - float(kl.double().sum())) / float(kl.double().sum())
        print("   KL batch total deviation:", tot)
        assert tot < tol, tot


@pytest.mark.parametrize("precision,gtol,gmean", [("tf32x3f", 5e-2, 2e-2), ("bf16x3f", 5e-2, 2e-2), ("bf16p3f", 5e-2, 2e-2), ("tf32", 5e-2, 2e-2),
                                                  ("fp32x3", 1e-3, None)])
def test_bench_shape_gradients_vs_oracle(precision, gtol, gmean):
    """Every parameter gradient of a B = 512 bench-shaped step (conditioned sigma) against the oracle's autograd."""
    cfg_o = om.Cfg(dec_type="lstm")
    p = _condition_sigma(cfg_o, om.init_params(cfg_o, seed=0))
    tokens, lens, classes, labels, eps = _bench_rows(512, 64, seed=21)
    eng = _engine(precision, p)
    out = eng.forward(_dev(tokens), _dev(lens), _dev(classes), _dev(labels), eps=_dev(eps, torch.float32))
    eng.backward()
    torch.cuda.synchronize()
    pp = {k: v.clone() for k, v in p.items()}
    opt = om.Adam(pp, clip_gradient=1.0)
    loss, ce, kl, probs, means, stds, grads = om.train_step(cfg_o, pp, opt, tokens, lens, classes, labels, eps)
    tol = FWD_TOL[precision]
    assert _rel(out["ce"], ce) < tol and _rel(out["kl"], kl) < tol and _rel(out["means"], means) < tol
    gmax = max(float(g.abs().max()) for g in grads.values())
    devs = []
    for n in eng.arena.names():
        scale = float(grads[n].abs().max())
        err = float((eng.arena.grad(n).cpu() - grads[n]).abs().max())
        if gmean is None:
            assert err <= gtol * scale + 2e-5 * gmax + 1e-7, (n, err, scale, gmax)
        elif scale > 1e-4 * gmax:
            devs.append((err / scale, n))
    if devs:
        devs.sort(reverse=True)
        print("%s gradient deviation, worst tensors:" % precision, devs[:4])
        assert devs[0][0] < gtol and sum(d for d, _ in devs) / len(devs) < gmean


# ------------------------------------------------------------------------------------------------ dropout
def _device_masks(cfg_o, B, T, p_drop, seed, sos_rows_only=True):
    """The keep masks the step's kernels draw for effective seed `seed`, keyed as oracle/model.py names its sites."""
    from musicstyletransfer_b200 import ops
    from musicstyletransfer_b200.engine import SITE_STRIDE
    D = cfg_o.enc_size
    masks = {}
    for l in range(cfg_o.enc_layers):
        prefix = "encoder.encoder.layer%d." % l
        top = sos_rows_only and l == cfg_o.enc_layers - 1
        for key, off, width in ((prefix + "att", 0, D), (prefix + ".ffh", 1, 4 * D), (prefix + "ff", 2, D)):
            if top:
                # the top encoder layer runs its row-wise part on the B SOS rows: the device draws a [B, width] mask for
                # them; every other position has no consumer (model.py:97-100), so its mask is irrelevant (ones)
                m = torch.empty(B * width, dtype=torch.uint8, device="cuda:0")
                ops.dropout_mask(m, p_drop, seed, l * SITE_STRIDE + off)
                full = torch.ones(B, T, width)
                full[:, 0, :] = m.view(B, width).float().cpu()
                masks[key] = full
            else:
                m = torch.empty(B * T * width, dtype=torch.uint8, device="cuda:0")
                ops.dropout_mask(m, p_drop, seed, l * SITE_STRIDE + off)
                masks[key] = m.view(B, T, width).float().cpu()
    return masks


def _device_eps(B, Z, seed):
    from musicstyletransfer_b200 import ops
    eps = torch.empty(B, Z, device="cuda:0")
    ops.normal_fill(eps, seed, 0xE95)
    return eps.cpu()


@pytest.mark.parametrize("precision,gtol", [("fp32", 1e-3), ("tf32x3f", 5e-2), ("bf16x3f", 5e-2), ("bf16p3f", 5e-2), ("tf32", 5e-2), ("fp32x3", 1e-3)])
def test_dropout_step_vs_oracle_with_device_masks(precision, gtol):
    """The train step as bench.py runs it (dropout 0.2): the oracle replays the step with the device's own keep masks
    (msx_dropout_mask) and eps (msx_normal_fill) -> losses, latent means and every gradient agree."""
    cfg_o = om.Cfg(dec_type="lstm", enc_dropout=0.2, dec_dropout=0.2)
    p = _condition_sigma(cfg_o, om.init_params(cfg_o, seed=0))
    B, T = 96, 65
    tokens, lens, classes, labels, _ = _bench_rows(B, 64, seed=31)
    eng = _engine(precision, p, dropout=0.2)
    out = eng.forward(_dev(tokens), _dev(lens), _dev(classes), _dev(labels), train=True)
    eng.backward()
    torch.cuda.synchronize()
    seed = BASE_SEED + 0                      # step 0 of a fresh engine
    masks = _device_masks(cfg_o, B, T, 0.2, seed)
    eps = _device_eps(B, 256, seed)
    keep = float(masks["encoder.encoder.layer0.att"].mean())
    assert 0.78 < keep < 0.82, keep
    pp = {k: v.clone() for k, v in p.items()}
    opt = om.Adam(pp, clip_gradient=1.0)
    loss, ce, kl, probs, means, stds, grads = om.train_step(cfg_o, pp, opt, tokens, lens, classes, labels, eps, masks=masks)
    dev = {"ce": _rel(out["ce"], ce), "kl": _rel(out["kl"], kl), "means": _rel(out["means"], means)}
    print("%s dropout-step forward deviation:" % precision, dev)
    assert max(dev.values()) < FWD_TOL.get(precision, 1e-4), dev
    gmax = max(float(g.abs().max()) for g in grads.values())
    worst = 0.0
    for n in eng.arena.names():
        scale = float(grads[n].abs().max())
        err = float((eng.arena.grad(n).cpu() - grads[n]).abs().max())
        if gtol <= 1e-3:
            assert err <= gtol * scale + 2e-5 * gmax + 1e-7, (n, err, scale, gmax)
        elif scale > 1e-4 * gmax:
            worst = max(worst, err / scale)
    assert worst < gtol, worst


def test_graph_replay_with_dropout_vs_oracle():
    """CUDA-graph replay of the dropout step (what bench.py times): with lr = 0 the parameters stay put, so replay s must
    equal the oracle's forward with the masks / eps of effective seed base + s (the device step counter)."""
    cfg_o = om.Cfg(dec_type="lstm", enc_dropout=0.2, dec_dropout=0.2)
    p = _condition_sigma(cfg_o, om.init_params(cfg_o, seed=0))
    B, T = 64, 65
    tokens, lens, classes, labels, _ = _bench_rows(B, 64, seed=41)
    eng = _engine("tf32x3f", p, dropout=0.2)
    args = [_dev(tokens), _dev(lens), _dev(classes), _dev(labels)]
    seen = []
    for s in range(4):                        # step 0 eager, step 1 capture + replay, steps 2.. replay
        out = eng.train_step_graphed(*args, kl_weight=1.0, global_batch=B, lr=0.0, clip_gradient=1.0)
        torch.cuda.synchronize()
        seen.append((out["ce"].clone().cpu(), out["kl"].clone().cpu(), out["means"].clone().cpu()))
    assert eng.step_count == 4 and int(eng.step_dev.item()) == 4
    for s in (0, 1, 3):
        masks = _device_masks(cfg_o, B, T, 0.2, BASE_SEED + s)
        eps = _device_eps(B, 256, BASE_SEED + s)
        with torch.no_grad():
            _, ce, kl, _, means, _ = om.step_losses(cfg_o, p, tokens, lens, classes, labels, eps, masks=masks)
        ce_d, kl_d, means_d = seen[s]
        dev = {"ce": _rel(ce_d, ce), "kl": _rel(kl_d, kl), "means": _rel(means_d, means)}
        print("replay %d deviation:" % s, dev)
        assert max(dev.values()) < FWD_TOL["tf32x3f"], (s, dev)
    # different steps draw different masks
    assert float((seen[1][0] - seen[2][0]).abs().max()) > 1e-3 * float(seen[1][0].abs().max())


# ------------------------------------------------------------------------------------------------ GEMM at the bench's M
STEP_SHAPES = [  # (N, K, transA, transB, what) at M = 2048 * 65 rows
    (768, 256, 0, 1, "QKV forward"), (1024, 256, 0, 1, "FF1 forward"), (256, 1024, 0, 1, "FF2 forward"),
    (1024, 256, 0, 0, "FF2 dgrad"), (256, 1024, 0, 0, "FF1 dgrad"), (256, 768, 0, 0, "QKV dgrad")]


@pytest.mark.parametrize("x3", [False, True])
@pytest.mark.parametrize("N,K,tA,tB,what", STEP_SHAPES)
def test_gemm_tc_at_bench_rows(N, K, tA, tB, what, x3):
    """msx_gemm_tc on the six forward / dgrad shapes of the step at the bench's M = 133 120 rows against a float64 matmul
    (TF32 operands: 2^-11 relative per operand; 3xTF32: fp32-level)."""
    from musicstyletransfer_b200 import ops
    M = 2048 * 65
    g = torch.Generator(device="cuda").manual_seed(N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    Bm = torch.randn((N, K) if tB else (K, N), device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    C = torch.empty(M, N, device="cuda")
    ops.gemm_tc(A, K, tA, Bm, K if tB else N, tB, C, N, M, N, K, bias=bias if tB else None, x3=x3)
    torch.cuda.synchronize()
    rows = torch.cat([torch.arange(0, 4096), torch.arange(M // 2, M // 2 + 2048), torch.arange(M - 4096, M)]).cuda()
    ref = A[rows].double() @ (Bm.double().t() if tB else Bm.double())
    if tB:
        ref += bias.double()
    err = float((C[rows].double() - ref).abs().max() / ref.abs().max())
    print("%s M=%d N=%d K=%d x3=%s: max err / max = %.3e" % (what, M, N, K, x3, err))
    # 3xTF32: what is left is fp32 accumulation order over K terms (measured 2e-6 .. 9e-6)
    assert err < (2e-5 if x3 else 1.5e-3), err
    assert torch.isfinite(C).all()


@pytest.mark.parametrize("N,K", [(768, 256), (1024, 256), (256, 1024)])
def test_gemm_tc_b3_at_bench_rows(N, K):
    """msx_gemm_tc_b3 (bf16 hi + lo operand split, three kind::f16 MMAs per k-step) on the step's forward shapes at the
    bench's M = 133 120 rows against float64: ~2^-17 per operand, 30x below single-pass TF32."""
    from musicstyletransfer_b200 import ops
    M = 2048 * 65
    g = torch.Generator(device="cuda").manual_seed(N + K + 1)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    C = torch.empty(M, N, device="cuda")
    ops.gemm_tc_b3(A, K, W, K, C, N, M, N, K, bias=bias)
    torch.cuda.synchronize()
    rows = torch.cat([torch.arange(0, 4096), torch.arange(M // 2, M // 2 + 2048), torch.arange(M - 4096, M)]).cuda()
    ref = A[rows].double() @ W.double().t() + bias.double()
    err = float((C[rows].double() - ref).abs().max() / ref.abs().max())
    print("bf16x3 M=%d N=%d K=%d: max err / max = %.3e" % (M, N, K, err))
    assert err < 5e-5, err
    assert torch.isfinite(C).all()


@pytest.mark.parametrize("M,N,K,relu,acc", [(37, 293, 132, False, False), (300, 64, 100, True, False), (2048, 256, 256, False, True),
                                            (1, 128, 32, False, False), (513, 1024, 160, True, False), (129, 320, 1024, False, False)])
def test_gemm_tc_b3_shapes_and_epilogues(M, N, K, relu, acc):
    """Odd numbers of 32-wide k-blocks (the zero-filled half stage), K % 32 != 0, M below one tile, N that pads, ReLU + bit
    mask, accumulate."""
    from musicstyletransfer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.1
    bias = torch.randn(N, device="cuda", generator=g)
    ldc = (N + 3) // 4 * 4
    C0 = torch.randn(M, ldc, device="cuda", generator=g)
    C = C0.clone()
    mask = torch.zeros(M, N // 32, dtype=torch.int32, device="cuda") if (relu and N % 32 == 0) else None
    ops.gemm_tc_b3(A, K, W, K, C, ldc, M, N, K, bias=None if acc else bias, relu=relu, accumulate=acc, mask_out=mask,
                   ldmask=N // 32)
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t()
    ref = ref + (C0[:, :N].double() if acc else bias.double())
    if relu:
        ref = ref.clamp(min=0)
    err = float((C[:, :N].double() - ref).abs().max() / ref.abs().max())
    assert err < 5e-5, err
    if mask is not None:
        bits = ((mask.view(M, N // 32, 1) >> torch.arange(32, device="cuda").view(1, 1, 32)) & 1).view(M, N).bool()
        want = ref > 0
        near = ref.abs() < 1e-4 * ref.abs().max()
        assert bool(((bits == want) | near).all())


@pytest.mark.parametrize("x3", [False, True])
def test_gemm_tc_wgrad_at_bench_rows(x3):
    """Split-K weight gradient dW += dY^T X with the reduction over the bench's 133 120 rows."""
    from musicstyletransfer_b200 import ops
    M = 2048 * 65
    g = torch.Generator(device="cuda").manual_seed(5)
    dY = torch.randn(M, 256, device="cuda", generator=g) * 0.1
    X = torch.randn(M, 1024, device="cuda", generator=g)
    gw = torch.zeros(256, 1024, device="cuda")
    ops.gemm_tc(dY, 256, 1, X, 1024, 0, gw, 1024, 256, 1024, M, splitk=ops.wgrad_splitk(256, 1024, M), x3=x3)
    torch.cuda.synchronize()
    ref = dY.double().t() @ X.double()
    err = float((gw.double() - ref).abs().max() / ref.abs().max())
    print("wgrad K=%d x3=%s: max err / max = %.3e" % (M, x3, err))
    assert err < (1e-4 if x3 else 2e-3), err          # fp32 accumulation over 133 120 terms: measured 2.6e-5


# ------------------------------------------------------------------------------------------------ LSTM recurrence
@pytest.mark.parametrize("B,T", [(64, 65), (37, 19), (256, 33), (1300, 9)])
def test_lstm_tc_vs_oracle_lstm_layer(B, T):
    """lstm_tc (TF32 mma.sync recurrence) directly against oracle.lstm_layer (gluon.rnn.LSTM restatement, model.py:148-153):
    hidden states forward; d(pre-activations), dh0 / dc0 and the bias gradients against its autograd."""
    from musicstyletransfer_b200 import ops
    H = 128
    g = torch.Generator().manual_seed(B + T)
    x = torch.randn(B, T, H, generator=g) * 0.5
    p = {"l0_i2h_weight": torch.randn(4 * H, H, generator=g) * 0.1, "l0_h2h_weight": torch.randn(4 * H, H, generator=g) * 0.1,
         "l0_i2h_bias": torch.randn(4 * H, generator=g) * 0.1, "l0_h2h_bias": torch.randn(4 * H, generator=g) * 0.1}
    tv = torch.randn(B, 2 * H, generator=g) * 0.5
    dhs = torch.randn(B, T, H, generator=g) * 0.3
    # oracle with autograd
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    tvo = tv.clone().requires_grad_(True)
    hs_o, _, _ = om.lstm_layer(x, tvo[:, :H], tvo[:, H:], po, "l0_")
    (hs_o * dhs).sum().backward()
    # device: i2h pre-activations (x W_i2h^T + b_i2h) computed exactly on the host, recurrence on the tensor-core kernel
    gx = (x.reshape(B * T, H) @ p["l0_i2h_weight"].t() + p["l0_i2h_bias"]).cuda().contiguous()
    w, bh, tvd = p["l0_h2h_weight"].cuda(), p["l0_h2h_bias"].cuda(), tv.cuda().contiguous()
    assert ops.lstm_tc_supported(H, 2 * H, tvd, tvd[:, H:])
    hs, hp, cs = (torch.zeros(B * T, H, device="cuda") for _ in range(3))
    ops.lstm_tc_fwd(gx, w, bh, tvd, tvd[:, H:], 2 * H, hs, hp, cs, B, T, H)
    dtv = torch.zeros(B, 2 * H, device="cuda")
    dbi, dbh = torch.zeros(4 * H, device="cuda"), torch.zeros(4 * H, device="cuda")
    ops.lstm_tc_bwd(gx, w, cs, tvd[:, H:], 2 * H, dhs.reshape(B * T, H).cuda().contiguous(), dtv, dtv[:, H:], B, T, H,
                    db_i2h=dbi, db_h2h=dbh)
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.cpu() - b).abs().max() / (b.abs().max() + 1e-12))
    dev = {"hs": rel(hs.view(B, T, H), hs_o.detach()), "dtv": rel(dtv, tvo.grad), "dbh": rel(dbh, po["l0_h2h_bias"].grad),
           "dbi": rel(dbi, po["l0_i2h_bias"].grad)}
    # weight gradient of the recurrent matrix from the kernel's d(pre-activations) and saved h_{t-1}
    dW = gx.view(B * T, 4 * H).t().double() @ hp.double()
    dev["dW_h2h"] = rel(dW.float(), po["l0_h2h_weight"].grad)
    print("lstm_tc vs oracle B=%d T=%d:" % (B, T), dev)
    assert dev["hs"] < 3e-3 and dev["dtv"] < 1e-2 and dev["dbh"] < 1e-2 and dev["dbi"] < 1e-2 and dev["dW_h2h"] < 1e-2, dev


def _planes(x):
    from musicstyletransfer_b200 import ops
    hi, lo = torch.empty_like(x, dtype=torch.bfloat16), torch.empty_like(x, dtype=torch.bfloat16)
    ops.split_planes(x, hi, lo)
    return hi, lo


def test_split_planes_carry_16_mantissa_bits():
    from musicstyletransfer_b200 import ops  # noqa: F401
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(1024, 256, device="cuda", generator=g) * torch.logspace(-6, 3, 256, device="cuda")
    hi, lo = _planes(x)
    assert torch.equal(hi, x.to(torch.bfloat16))
    assert torch.equal(lo, (x - hi.float()).to(torch.bfloat16))
    rel = ((hi.double() + lo.double() - x.double()).abs() / x.double().abs().clamp(min=1e-30)).max()
    assert float(rel) < 2.0 ** -16, float(rel)


@pytest.mark.parametrize("N,K", [(768, 256), (1024, 256), (256, 1024)])
def test_gemm_tc_p3_at_bench_rows(N, K):
    """msx_gemm_tc_p3 (operands as bf16 hi / lo planes, hi*hi + hi*lo + lo*hi walks on kind::f16) on the step's forward
    shapes at the bench's M = 133 120 rows against float64."""
    from musicstyletransfer_b200 import ops
    M = 2048 * 65
    g = torch.Generator(device="cuda").manual_seed(N + K + 1)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    C = torch.empty(M, N, device="cuda")
    (Ah, Al), (Wh, Wl) = _planes(A), _planes(W)
    assert ops.gemm_tc_p3_supported(Ah, K, Wh, K, C, N, M, N, K)
    ops.gemm_tc_p3(Ah, Al, K, Wh, Wl, K, C, N, M, N, K, bias=bias)
    torch.cuda.synchronize()
    rows = torch.cat([torch.arange(0, 4096), torch.arange(M // 2, M // 2 + 2048), torch.arange(M - 4096, M)]).cuda()
    ref = A[rows].double() @ W.double().t() + bias.double()
    err = float((C[rows].double() - ref).abs().max() / ref.abs().max())
    print("bf16p3 M=%d N=%d K=%d: max err / max = %.3e" % (M, N, K, err))
    assert err < 2e-5, err
    assert torch.isfinite(C).all()


@pytest.mark.parametrize("M,N,K,relu,acc,cplanes", [(37, 293, 128, False, False, False), (300, 64, 64, True, False, False),
                                                    (2048, 256, 256, False, True, False), (1, 128, 64, False, False, False),
                                                    (513, 1024, 192, True, False, True), (129, 320, 1024, False, False, True),
                                                    (4160, 1024, 256, True, False, True)])
def test_gemm_tc_p3_shapes_and_epilogues(M, N, K, relu, acc, cplanes):
    """1-CTA and pair kernels, M below one tile, N that pads, ReLU + dropout-free bit mask, accumulate, result as planes."""
    from musicstyletransfer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.1
    bias = torch.randn(N, device="cuda", generator=g)
    ldc = (N + 7) // 8 * 8
    C0 = torch.randn(M, ldc, device="cuda", generator=g)
    (Ah, Al), (Wh, Wl) = _planes(A), _planes(W)
    mask = torch.zeros(M, N // 32, dtype=torch.int32, device="cuda") if (relu and N % 32 == 0) else None
    if cplanes:
        Ch = torch.zeros(M, ldc, device="cuda", dtype=torch.bfloat16)
        Cl = torch.zeros(M, ldc, device="cuda", dtype=torch.bfloat16)
        ops.gemm_tc_p3(Ah, Al, K, Wh, Wl, K, Ch, ldc, M, N, K, bias=bias, relu=relu, mask_out=mask, ldmask=N // 32, C_lo=Cl)
        C = Ch.float() + Cl.float()
    else:
        C = C0.clone()
        ops.gemm_tc_p3(Ah, Al, K, Wh, Wl, K, C, ldc, M, N, K, bias=None if acc else bias, relu=relu, accumulate=acc,
                       mask_out=mask, ldmask=N // 32)
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t()
    ref = ref + (C0[:, :N].double() if acc else bias.double())
    if relu:
        ref = ref.clamp(min=0)
    err = float((C[:, :N].double() - ref).abs().max() / ref.abs().max())
    assert err < 3e-5, err
    if mask is not None:
        bits = ((mask.view(M, N // 32, 1) >> torch.arange(32, device="cuda").view(1, 1, 32)) & 1).view(M, N).bool()
        want = ref > 0
        near = ref.abs() < 1e-4 * ref.abs().max()
        assert bool(((bits == want) | near).all())
