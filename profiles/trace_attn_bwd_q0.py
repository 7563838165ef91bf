"""clock64 trace of the pipelined attention backward (general and q0_only): per item of block 0, cycles from the loop top to
1 tiles landed, 2 MMA S(/dP) issued, 3 p_ready seen by the issuer, 4 softmax done, 6 output MMAs issued, 7 outputs ready,
8 stores issued (stamps of msx_attention_tc_set_trace)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from musicstyletransfer_b200 import ops, lib
dev = "cuda"
B, T, H, dh = 2048, 65, 8, 32
qkv = torch.randn(B * T, 3 * H * dh, device=dev); mask = torch.ones(B * T, device=dev)
dctx = torch.randn(B * T, H * dh, device=dev); dqkv = torch.empty_like(qkv); db = torch.zeros(3 * H * dh, device=dev)
d0 = torch.zeros(B, T, H * dh, device=dev); d0[:, 0] = 1.0; d0 = d0.view(B * T, H * dh)
tr = torch.zeros(2 * 16 * 9, dtype=torch.int64, device=dev)
L = lib.load()
for q0 in (False, True):
    dd = d0 if q0 else dctx
    for _ in range(2):
        ops.attention_tc_bwd(qkv, mask, dd, dqkv, B, T, H, dh, dbias=db, q0_only=q0)
    L.msx_attention_tc_set_trace(ctypes.c_void_p(tr.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.attention_tc_bwd(qkv, mask, dd, dqkv, B, T, H, dh, dbias=db, q0_only=q0); e1.record()
    torch.cuda.synchronize()
    L.msx_attention_tc_set_trace(None)
    print("q0_only", q0, "kernel ms", e0.elapsed_time(e1))
    t = tr.cpu().view(2, 16, 9)
    for g in range(2):
        print(" group", g)
        for n in range(2, 8):
            r = t[g, n]; base = int(r[0])
            print("  item", n, "top@%d" % (int(r[0]) - int(t[g, 2, 0])), [int(r[i]) - base for i in range(1, 9)])
