"""Worst per-tensor gradient deviations of the long-row step (T = 129 / 257) vs the oracle for a few seeds and precisions:
separates a kernel error (every seed, every precision) from ReLU-mask flips of the 8-row top-layer feed-forward under
single-pass TF32 (one tensor, one seed).  Run on the GPU box: python profiles/micro/diag_long_rows.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
import test_engine_gpu as te
from musicstyletransfer_b200.engine import VAEConfig, VAEEngine
om = te.om

for T in (129, 257):
    for prec in ("tf32", "bf16p3f"):
        for seed in (T, T + 1, T + 2):
            cfg_o = om.Cfg(dec_type="lstm")
            p = te._condition_sigma(cfg_o, om.init_params(cfg_o, seed=0))
            tokens, seq_lens, classes, labels, eps = te._batch(8, T, 293, 2, 256, seed=seed, min_len=T // 2)
            eng = VAEEngine(VAEConfig(dec_type="lstm"), "cuda:0", precision=prec)
            eng.arena.load_state(p)
            out = eng.forward(te._dev(tokens), te._dev(seq_lens), te._dev(classes), te._dev(labels), eps=te._dev(eps, torch.float32))
            opt = om.Adam({k: v.clone() for k, v in p.items()}, clip_gradient=1.0)
            pp = {k: v.clone() for k, v in p.items()}
            loss, ce, kl, probs, means, stds, grads = om.train_step(cfg_o, pp, opt, tokens, seq_lens, classes, labels, eps)
            rel = lambda a, b: float((a.cpu() - b).abs().max() / b.abs().max())
            eng.backward()
            torch.cuda.synchronize()
            gmax = max(float(g.abs().max()) for g in grads.values())
            devs = sorted(((rel(eng.arena.grad(n), grads[n]), n) for n in eng.arena.names()
                           if float(grads[n].abs().max()) > 1e-4 * gmax), reverse=True)
            print("T=%d %s seed=%d means %.1e  worst: %s  mean %.2e" % (T, prec, seed, rel(out["means"], means),
                  ", ".join("%s %.1e" % (n.replace("encoder.encoder.", "enc."), d) for d, n in devs[:3]), sum(d for d, _ in devs) / len(devs)))
