"""Tensor-core LSTM recurrence (TF32 mma.sync, lstm_tc.cu) vs the exact-fp32 kernel (lstm.cu), which the engine tests
pin to the oracle: forward states / gate activations and backward d(pre-activations), dh0, dc0, bias gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(B, T, seed):
    H = 128
    g = torch.Generator().manual_seed(seed)
    gx = (torch.randn(B * T, 4 * H, generator=g) * 0.8).cuda()
    w = (torch.randn(4 * H, H, generator=g) * 0.12).cuda()
    bh = (torch.randn(4 * H, generator=g) * 0.1).cuda()
    tv = (torch.randn(B, 2 * H, generator=g) * 0.5).cuda()
    dhs = (torch.randn(B * T, H, generator=g) * 0.3).cuda()
    return H, gx, w, bh, tv, dhs


# B <= 16 * 74 runs the 16-rows-per-cluster kernels (all clusters resident), larger batches the 32-row kernels (forward:
# two sub-blocks with the warp groups one phase apart, persistent over row blocks)
@pytest.mark.parametrize("B,T", [(32, 5), (64, 65), (37, 19), (1, 3), (300, 8), (1200, 4), (1500, 33), (2500, 7), (5000, 1)])
def test_lstm_tc_matches_fp32_kernel(B, T):
    from musicstyletransfer_b200 import ops
    H, gx, w, bh, tv, dhs = _case(B, T, seed=B * 100 + T)
    assert ops.lstm_tc_supported(H, 2 * H, tv, tv[:, H:])
    out = {}
    for name, fwd, bwd in (("f32", ops.lstm_fwd, ops.lstm_bwd), ("tc", ops.lstm_tc_fwd, ops.lstm_tc_bwd)):
        gates = gx.clone()
        hs, hp, cs = (torch.zeros(B * T, H, device="cuda") for _ in range(3))
        fwd(gates, w, bh, tv, tv[:, H:], 2 * H, hs, hp, cs, B, T, H)
        act = gates.clone()
        dtv = torch.zeros(B, 2 * H, device="cuda")
        dbi, dbh = torch.zeros(4 * H, device="cuda"), torch.zeros(4 * H, device="cuda")
        bwd(gates, w, cs, tv[:, H:], 2 * H, dhs, dtv, dtv[:, H:], B, T, H, db_i2h=dbi, db_h2h=dbh)
        torch.cuda.synchronize()
        out[name] = dict(act=act, hs=hs, hp=hp, cs=cs, dgates=gates, dtv=dtv, dbi=dbi, dbh=dbh)
    # TF32 operands (10-bit mantissa) on a 128-long dot product, fp32 accumulation: errors ~1e-3 of the scale,
    # compounding mildly through the T-step recurrence
    for k, tol in (("act", 3e-3), ("hs", 3e-3), ("hp", 3e-3), ("cs", 3e-3), ("dgates", 1e-2), ("dtv", 1e-2), ("dbi", 1e-2),
                   ("dbh", 1e-2)):
        a, b = out["tc"][k], out["f32"][k]
        err = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-9)
        assert err < tol, (k, err)
    assert torch.allclose(out["tc"]["dbi"], out["tc"]["dbh"], rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------------ any hidden size
# --d-hidden is a free flag of the reference's CLI (LSTMConfig.hidden_dim, model.py:131-153): sizes other than 32 / 64 / 128
# take the L2-streaming kernels of lstm.cu (lstm_gen_*), exact fp32 like the resident ones.
@pytest.mark.parametrize("B,T,H", [(19, 7, 48), (64, 33, 100), (9, 5, 1), (130, 12, 200), (33, 9, 257), (16, 4, 512)])
def test_lstm_any_hidden_size_vs_oracle(B, T, H):
    from musicstyletransfer_b200 import ops
    from oracle import model as om
    g = torch.Generator().manual_seed(B * 1000 + T * 10 + H)
    X = 24
    x = torch.randn(B, T, X, generator=g) * 0.5
    p = {"l0_i2h_weight": torch.randn(4 * H, X, generator=g) * 0.2, "l0_h2h_weight": torch.randn(4 * H, H, generator=g) * (0.8 / H ** 0.5),
         "l0_i2h_bias": torch.randn(4 * H, generator=g) * 0.1, "l0_h2h_bias": torch.randn(4 * H, generator=g) * 0.1}
    tv = torch.randn(B, 2 * H, generator=g) * 0.5
    dhs = torch.randn(B, T, H, generator=g) * 0.3
    po = {k: v.double().requires_grad_(True) for k, v in p.items()}
    tvo = tv.double().requires_grad_(True)
    hs_o, _, c_o = om.lstm_layer(x.double(), tvo[:, :H], tvo[:, H:], po, "l0_")
    (hs_o * dhs.double()).sum().backward()
    gx = (x.reshape(B * T, X).double() @ p["l0_i2h_weight"].double().t() + p["l0_i2h_bias"].double()).float().cuda().contiguous()
    w, bh, tvd = p["l0_h2h_weight"].cuda(), p["l0_h2h_bias"].cuda(), tv.cuda().contiguous()
    hs, hp, cs = (torch.zeros(B * T, H, device="cuda") for _ in range(3))
    ops.lstm_fwd(gx, w, bh, tvd, tvd[:, H:], 2 * H, hs, hp, cs, B, T, H)
    dtv = torch.zeros(B, 2 * H, device="cuda")
    dbi, dbh = torch.zeros(4 * H, device="cuda"), torch.zeros(4 * H, device="cuda")
    ops.lstm_bwd(gx, w, cs, tvd[:, H:], 2 * H, dhs.reshape(B * T, H).cuda().contiguous(), dtv, dtv[:, H:], B, T, H,
                 db_i2h=dbi, db_h2h=dbh)
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double().cpu() - b).abs().max() / (b.abs().max() + 1e-12))
    dW = gx.view(B * T, 4 * H).t().double() @ hp.double()
    dev = {"hs": rel(hs.view(B, T, H), hs_o.detach()), "c_T": rel(cs.view(B, T, H)[:, -1], c_o.detach()),
           "hprev0": rel(hp.view(B, T, H)[:, 0], tv[:, :H].double()), "dtv": rel(dtv, tvo.grad),
           "dbh": rel(dbh, po["l0_h2h_bias"].grad), "dbi": rel(dbi, po["l0_i2h_bias"].grad),
           "dW_h2h": rel(dW, po["l0_h2h_weight"].grad)}
    assert max(dev.values()) < 1e-4, dev             # fp32 FMA chains + fast-math exp / tanh vs float64: measured <= 2e-5


def test_lstm_any_hidden_size_rejects_oversize():
    from musicstyletransfer_b200 import lib, ops
    H, B, T = 520, 2, 2
    z = lambda *s: torch.zeros(*s, device="cuda")
    with pytest.raises(lib.MsxError):
        ops.lstm_fwd(z(B * T, 4 * H), z(4 * H, H), z(4 * H), z(B, 2 * H), z(B, 2 * H)[:, H:], 2 * H, z(B * T, H), z(B * T, H),
                     z(B * T, H), B, T, H)
