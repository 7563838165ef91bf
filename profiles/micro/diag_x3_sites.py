"""Which forward GEMMs need the 3xTF32 compensation?  Latent-means deviation (worst of B = 2048 / 2 x B = 512, raw weights) of
the tf32x3f mode with ONE group of encoder GEMMs left single-pass TF32.   python profiles/micro/diag_x3_sites.py"""
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
for B, seed in ((2048, 0), (512, 1), (512, 2)):
    p = om.init_params(cfg_o, seed=seed)
    tok, lens, cls, lab = synth.token_rows_4_4(B, 64, seed=10 + seed)
    eps = torch.randn(B, 256, generator=torch.Generator().manual_seed(10 + seed))
    with torch.no_grad():
        _, ce, kl, _, means, stds = om.step_losses(cfg_o, p, f(tok), f(lens), f(cls), f(lab), eps)
    cases.append((p, (tok, lens, cls, lab), eps, means))
groups = [(), ("layer0.qkv",), ("layer0.self_attention.W_proj",), ("layer0.ff.ff1",), ("layer0.ff.ff2",), ("layer1.qkv",),
          ("layer0.ff.ff1", "layer0.ff.ff2"), ("layer0.qkv", "layer1.qkv"), ("layer1.self_attention", "layer1.ff", "latent_proj"),
          ("layer0", "layer1", "latent_proj")]
for skip in groups:
    worst = 0.0
    for p, (tok, lens, cls, lab), eps, means in cases:
        eng = VAEEngine(VAEConfig(dec_type="lstm"), dev, precision="tf32x3f")
        eng.x3_skip = skip
        eng.arena.load_state(p)
        out = eng.forward(t(tok), t(lens), t(cls), t(lab), eps=eps.to(dev))
        torch.cuda.synchronize()
        worst = max(worst, rel(out["means"], means))
        del eng
    print("single-pass TF32: %-70s latent means dev %.2e" % (", ".join(skip) or "(none: full tf32x3f)", worst), flush=True)
