"""Per-launch time of the tensor-core LSTM recurrence (lstm_tc.cu) at the bench shapes: B = 2048 / 32, T = 65, H = 128.
Run on the GPU box: python profiles/micro/lstm_rate.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from musicstyletransfer_b200 import ops

H = 128
for B, T in ((2048, 65), (32, 65), (256, 65), (1024, 65), (8192, 1)):
    g = torch.Generator().manual_seed(1)
    gx = (torch.randn(B * T, 4 * H, generator=g) * 0.8).cuda()
    w = (torch.randn(4 * H, H, generator=g) * 0.12).cuda()
    bh = (torch.randn(4 * H, generator=g) * 0.1).cuda()
    tv = (torch.randn(B, 2 * H, generator=g) * 0.5).cuda()
    dhs = (torch.randn(B * T, H, generator=g) * 0.3).cuda()
    hs, hp, cs = (torch.zeros(B * T, H, device="cuda") for _ in range(3))
    dtv = torch.zeros(B, 2 * H, device="cuda")
    dbi, dbh = torch.zeros(4 * H, device="cuda"), torch.zeros(4 * H, device="cuda")
    gates = gx.clone()
    def fwd():
        ops.lstm_tc_fwd(gates, w, bh, tv, tv[:, H:], 2 * H, hs, hp, cs, B, T, H)
    def bwd():
        ops.lstm_tc_bwd(gates, w, cs, tv[:, H:], 2 * H, dhs, dtv, dtv[:, H:], B, T, H, db_i2h=dbi, db_h2h=dbh)
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / n
        print("lstm_tc_%s B=%d T=%d: %.1f us per launch, %.2f us per step" % (name, B, T, us, us / T))
