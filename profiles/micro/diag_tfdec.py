import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from oracle import model as om
import test_engine_gpu as t
from musicstyletransfer_b200.engine import VAEConfig, VAEEngine
cfg_o = om.Cfg(dec_type="transformer")
p = t._condition_sigma(cfg_o, om.init_params(cfg_o, seed=0))
tokens, seq_lens, classes, labels, eps = t._batch(64, 65, 293, 2, 256, seed=1, min_len=33)
opt = om.Adam({k: v.clone() for k, v in p.items()}, clip_gradient=1.0)
pp = {k: v.clone() for k, v in p.items()}
loss, ce, kl, probs, means, stds, grads = om.train_step(cfg_o, pp, opt, tokens, seq_lens, classes, labels, eps)
gmax = max(float(g.abs().max()) for g in grads.values())
res = {}
for prec in ("tf32", "bf16"):
    eng = VAEEngine(VAEConfig(dec_type="transformer"), "cuda:0", precision=prec)
    eng.arena.load_state(p)
    eng.forward(t._dev(tokens), t._dev(seq_lens), t._dev(classes), t._dev(labels), eps=t._dev(eps, torch.float32))
    eng.backward(); torch.cuda.synchronize()
    res[prec] = {n: eng.arena.grad(n).cpu().clone() for n in eng.arena.names()}
print("gmax", gmax)
for n in grads:
    if not n.startswith("decoder"): continue
    s = float(grads[n].abs().max())
    e1 = float((res["tf32"][n] - grads[n]).abs().max()); e2 = float((res["bf16"][n] - grads[n]).abs().max())
    print("%-50s scale %.3e (%.1e of gmax)  tf32 err %.2e (%.3f)  bf16 err %.2e (%.3f)" % (n, s, s / gmax, e1, e1 / max(s, 1e-30), e2, e2 / max(s, 1e-30)))
