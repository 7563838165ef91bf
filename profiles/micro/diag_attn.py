"""Diagnostic: tcgen05 attention backward vs autograd for several shapes / head widths / output types."""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import test_attention_gpu as t
from musicstyletransfer_b200 import ops
for (B, T, H, dh) in [(3, 65, 8, 32), (2, 16, 4, 32), (4, 97, 3, 32), (3, 66, 8, 16), (4, 97, 2, 16)]:
    for odt in (torch.float32, torch.bfloat16):
        qkv, mask = t._inputs(B, T, H, dh, seed=100 + T)
        g = torch.Generator().manual_seed(T)
        dctx = torch.randn(B * T, H * dh, generator=g)
        x = qkv.double().requires_grad_(True)
        (t._ref_fwd(x, mask, B, T, H, dh) * dctx.double()).sum().backward()
        want = x.grad
        out = torch.full((B * T, 3 * H * dh), 5.0, device="cuda", dtype=odt)
        db = torch.zeros(3 * H * dh, device="cuda")
        ops.attention_tc_bwd(qkv.cuda(), mask.cuda(), dctx.cuda(), out, B, T, H, dh, dbias=db)
        torch.cuda.synchronize()
        d = (out.double().cpu() - want).abs()
        D = H * dh
        print(B, T, H, dh, odt, "err", float(d.max() / want.abs().max()), "parts K/Q/V", [float(d[:, i * D:(i + 1) * D].max()) for i in range(3)],
              "dbias err", float((db.double().cpu() - want.sum(0)).abs().max()))
