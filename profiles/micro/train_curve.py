"""Loss curve of the graph-replayed train step on the synthetic 4/4 rows (B = 2048, L = 64, train-vae.sh model, dropout 0.2,
Adam 3e-4, clip 1.0) for the TF32 path and the bf16 variant: the step trains, and the two precisions follow each other."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from musicstyletransfer_b200 import synth
from musicstyletransfer_b200.engine import VAEConfig, VAEEngine
dev = torch.device("cuda", 0)
B, L, nb = 2048, 64, 16
tok, lens, cls, lab = synth.token_rows_4_4(B * nb, L, seed=100)
bat = [tuple(torch.from_numpy(a[i * B:(i + 1) * B].copy()).to(dev) for a in (tok, lens, cls, lab)) for i in range(nb)]
for prec in ("tf32", "bf16"):
    eng = VAEEngine(VAEConfig(dec_type="lstm", enc_dropout=0.2, dec_dropout=0.2), dev, seed=0, precision=prec)
    line = []
    for step in range(401):
        out = eng.train_step_graphed(*bat[step % nb], kl_weight=1.0, global_batch=B, lr=3e-4, clip_gradient=1.0)
        if step % 50 == 0:
            line.append("%d: ce %.3f kl %.2f" % (step, float(out["ce"].mean()), float(out["kl"].mean())))
    print(prec, " | ".join(line))
