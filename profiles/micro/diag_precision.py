"""Which precision mode meets the north star's 1e-3 on loss / KL / latent means at the bench shape, and what does it cost?
   python profiles/micro/diag_precision.py            (on the GPU box)
Prints, per mode, the worst forward deviation vs the fp32 oracle over B = 2048 (seed 0) and B = 512 (seeds 1-3), raw Xavier
weights, and the graph-replayed step time at B = 2048."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import model as om                                     # noqa: E402
from musicstyletransfer_b200 import synth                          # noqa: E402
from musicstyletransfer_b200.engine import VAEConfig, VAEEngine    # noqa: E402

dev = "cuda:0"
t = lambda a: torch.from_numpy(a).to(dev)
f = lambda a: torch.from_numpy(a).float()
rel = lambda a, b: float((a.float().cpu() - b).abs().max() / b.abs().max())
cfg_o = om.Cfg(dec_type="lstm")
cases = []
for B, seed in ((2048, 0), (512, 1), (512, 2), (512, 3)):
    p = om.init_params(cfg_o, seed=seed)
    tok, lens, cls, lab = synth.token_rows_4_4(B, 64, seed=10 + seed)
    eps = torch.randn(B, 256, generator=torch.Generator().manual_seed(10 + seed))
    with torch.no_grad():
        _, ce, kl, _, means, stds = om.step_losses(cfg_o, p, f(tok), f(lens), f(cls), f(lab), eps)
    cases.append((B, seed, p, (tok, lens, cls, lab), eps, ce, kl, means, stds))

modes = [("tf32", {}), ("tf32x3f", {}), ("bf16x3f", {}), ("tf32x3f + exact attention", {"attn_tc": False}), ("fp32x3", {}),
         ("fp32x3 + tc attention/lstm", {"attn_tc": True, "lstm_tc": True}), ("fp32", {})]
for name, over in modes:
    prec = name.split(" ")[0]
    worst = {"ce": 0.0, "means": 0.0, "stds": 0.0, "kl_total": 0.0}
    for B, seed, p, (tok, lens, cls, lab), eps, ce, kl, means, stds in cases:
        eng = VAEEngine(VAEConfig(dec_type="lstm"), dev, precision=prec)
        for k, v in over.items():
            setattr(eng, k, v)
        eng.arena.load_state(p)
        out = eng.forward(t(tok), t(lens), t(cls), t(lab), eps=eps.to(dev))
        torch.cuda.synchronize()
        worst["ce"] = max(worst["ce"], rel(out["ce"], ce))
        worst["means"] = max(worst["means"], rel(out["means"], means))
        worst["stds"] = max(worst["stds"], rel(out["stds"], stds))
        worst["kl_total"] = max(worst["kl_total"], abs(float(out["kl"].double().sum().cpu()) - float(kl.double().sum())) / float(kl.double().sum()))
        del eng
    # step time, dropout 0.2, graph replay
    eng = VAEEngine(VAEConfig(dec_type="lstm", enc_dropout=0.2, dec_dropout=0.2), dev, precision=prec)
    for k, v in over.items():
        setattr(eng, k, v)
    tok, lens, cls, lab = synth.token_rows_4_4(2048, 64, seed=5)
    args = [t(tok), t(lens), t(cls), t(lab)]
    for _ in range(4):
        eng.train_step_graphed(*args, global_batch=2048, clip_gradient=1.0)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        eng.train_step_graphed(*args, global_batch=2048, clip_gradient=1.0)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print("%-32s worst dev: ce %.2e means %.2e stds %.2e kl_total %.2e | %.3f ms/step = %.0f seq/s" % (
        name, worst["ce"], worst["means"], worst["stds"], worst["kl_total"], ms, 2048 / ms * 1e3), flush=True)
    eng._graphs.clear()
    del eng
