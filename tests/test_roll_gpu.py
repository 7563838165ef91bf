"""Piano-roll training path (`--featurisation roll`): K1 roll -> msx_roll_features -> Dense-embedded encoder / LSTM decoder
-> sigmoid BCE (loss.py:27-81) + KL, against the oracle's restatement (oracle/roll_model.py; BCE itself is pinned to the
reference-generated loss goldens).  Tolerance 1e-3 relative on bce / KL / latent means, gradients 1e-3 of their scale on the
exact path."""
import numpy as np
import pytest
import torch

from oracle import model as om
from oracle import roll_model as rm

pytestmark = pytest.mark.gpu


def _roll(B, S, seed, density=0.04):
    g = torch.Generator().manual_seed(seed)
    roll = (torch.rand(B, S, 128, generator=g) < density)
    roll[0, :, :] = False                                  # an empty window (n_pos = 0 in the down-weighting)
    roll = (roll * torch.randint(1, 128, (B, S, 128), generator=g)).to(torch.uint8)   # velocity roll: label = (v > 0)
    classes = torch.randint(0, 2, (B,), generator=g)
    return roll, classes


def _engine(cfg_o, p, precision, dropout=0.0):
    from musicstyletransfer_b200.engine import VAEConfig, VAEEngine
    cfg = VAEConfig(num_classes=cfg_o.num_classes, enc_size=cfg_o.enc_size, enc_layers=cfg_o.enc_layers,
                    enc_heads=cfg_o.enc_heads, latent=cfg_o.latent, dec_type="lstm", dec_size=cfg_o.dec_size,
                    enc_dropout=dropout, dec_dropout=dropout, featurisation="roll")
    eng = VAEEngine(cfg, "cuda:0", precision=precision)
    assert set(eng.arena.names()) == set(p)
    eng.arena.load_state(p)
    return eng


def _conditioned(cfg_o, seed):
    p = rm.init_params(cfg_o, seed=seed)
    Z = cfg_o.latent
    p["encoder.latent_proj.weight"][Z:] *= 0.05
    p["encoder.latent_proj.bias"][Z:] = 3.0
    g = torch.Generator().manual_seed(seed + 1)
    for k in p:                                            # non-trivial biases so every gradient path is exercised
        if k.endswith("bias") and "latent_proj" not in k:
            p[k] = 0.05 * torch.randn(p[k].shape, generator=g)
    return p


def test_roll_features_match_oracle():
    from musicstyletransfer_b200 import ops
    roll, _ = _roll(5, 7, seed=1, density=0.2)
    renc = torch.full((5 * 8, 132), 9.0, device="cuda")
    rdec = torch.full((5 * 7, 132), 9.0, device="cuda")
    ops.roll_features(roll.cuda(), renc, rdec, 5, 7)
    want_e, want_d = rm.roll_features(roll)
    assert torch.equal(renc.cpu().view(5, 8, 132), want_e) and torch.equal(rdec.cpu().view(5, 7, 132), want_d)


@pytest.mark.parametrize("precision,ftol,gtol", [("fp32", 1e-4, 1e-3), ("fp32x3", 1e-4, 1e-3), ("tf32x3f", 3e-4, 5e-2), ("bf16p3f", 3e-4, 5e-2)])
@pytest.mark.parametrize("smoothing,downweight", [(0.0, True), (0.1, False)])
def test_roll_step_vs_oracle(precision, ftol, gtol, smoothing, downweight):
    cfg_o = om.Cfg(dec_type="lstm")                        # scripts/train-vae.sh sizes: enc 2x256/8h, Z=256, dec 1x128
    p = _conditioned(cfg_o, seed=3)
    B, S = 24, 64
    roll, classes = _roll(B, S, seed=5)
    eps = torch.randn(B, 256, generator=torch.Generator().manual_seed(9))
    eng = _engine(cfg_o, p, precision)
    out = eng.forward_roll(roll.cuda(), classes.to(torch.int32).cuda(), eps=eps.cuda(), label_smoothing=smoothing,
                           downweight=downweight)
    eng.backward_roll(kl_weight=1.0)
    torch.cuda.synchronize()
    pp = {k: v.clone() for k, v in p.items()}
    loss, bce, kl, logits, means, stds, grads = rm.train_step(cfg_o, pp, om.Adam(pp), roll, classes.float(), eps,
                                                              label_smoothing=smoothing, downweight=downweight)
    rel = lambda a, b: float((a.float().cpu() - b).abs().max() / b.abs().max())
    dev = {"bce": rel(out["bce"], bce), "kl": rel(out["kl"], kl), "means": rel(out["means"], means),
           "logits": rel(out["logits"], logits)}
    print("roll step %s smoothing=%g downweight=%s:" % (precision, smoothing, downweight), dev)
    assert dev["bce"] < ftol * 10 and dev["kl"] < ftol and dev["means"] < ftol, dev
    gmax = max(float(g.abs().max()) for g in grads.values())
    worst = 0.0
    for n in eng.arena.names():
        scale = float(grads[n].abs().max())
        err = float((eng.arena.grad(n).cpu() - grads[n]).abs().max())
        if gtol <= 1e-3:
            assert err <= gtol * scale + 2e-5 * gmax + 1e-7, (n, err, scale, gmax)
        elif scale > 1e-4 * gmax:
            worst = max(worst, err / scale)
    assert worst < gtol, worst


def test_roll_pipeline_from_note_events_trains():
    """note events -> K1 (binary roll) -> roll step, graph replayed: the reconstruction loss goes down."""
    from musicstyletransfer_b200 import featurise, synth
    from musicstyletransfer_b200.engine import VAEConfig, VAEEngine
    dtick, pitch, vel, offs = synth.note_events(n_seq=64, ev_per_seq=32, seed=2)
    t = lambda a: torch.from_numpy(a).cuda()
    _, roll, _ = featurise.rasterize(t(dtick), t(pitch), t(vel), t(offs))
    classes = torch.from_numpy((np.arange(64) % 2).astype(np.int32)).cuda()
    eng = VAEEngine(VAEConfig(dec_type="lstm", featurisation="roll", enc_dropout=0.1, dec_dropout=0.1), "cuda:0", seed=1,
                    precision="tf32x3f")
    hist = []
    for _ in range(40):
        out = eng.train_step_roll_graphed(roll, classes, global_batch=64, lr=1e-3, clip_gradient=1.0)
        hist.append(float(out["bce"].mean()))
    torch.cuda.synchronize()
    print("roll bce first/last:", hist[0], hist[-1])
    assert np.isfinite(hist).all() and hist[-1] < 0.9 * hist[0]
