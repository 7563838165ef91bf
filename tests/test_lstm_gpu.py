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
