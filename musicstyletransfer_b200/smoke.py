"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0, checked against the oracle."""
import numpy as np
import torch


def run():
    from oracle import featurise as of
    from oracle import model as om
    from . import featurise, synth
    from .engine import VAEConfig, VAEEngine

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    dev = "cuda:0"
    # K1: rasterise 256 sequences, bit-exact vs the oracle
    dtick, pitch, vel, offs = synth.note_events(n_seq=256, ev_per_seq=32, seed=0)
    t = lambda a: torch.from_numpy(a).to(dev)
    tok, roll, cnt = featurise.rasterize(t(dtick), t(pitch), t(vel), t(offs))
    otok, oroll, ocnt = of.rasterize_batch(dtick, pitch, vel, offs)
    assert np.array_equal(tok.cpu().numpy(), otok) and np.array_equal(roll.cpu().numpy(), oroll)
    assert np.array_equal(cnt.cpu().numpy(), ocnt)
    # one train step of the scripts/train-vae.sh model (enc 2x256 / 8 heads, Z = 256, LSTM decoder 1x128; B = 64, T = 65)
    # on the paths the bench times: the headline mode "bf16p3f" (gemm_tc2 / gemm_tc on kind::f16 planes forward and
    # kind::tf32 backward, attn_tc_* with compensated scores, lstm_tc_*) and the strict-fp32 "fp32x3" mode (gemm_tc2x3),
    # both against the fp32 oracle at the north star's 1e-3
    cfg_o = om.Cfg(dec_type="lstm")
    params = om.init_params(cfg_o, seed=0)
    Z = cfg_o.latent
    params["encoder.latent_proj.weight"][Z:] *= 0.05        # sigma away from 0 (KL holds log sigma^2)
    params["encoder.latent_proj.bias"][Z:] = 3.0
    tokens, lens, classes, labels = synth.token_rows_4_4(64, 64, seed=0)
    eps = torch.randn(64, Z, generator=torch.Generator().manual_seed(0))
    f = lambda a: torch.from_numpy(a).float()
    _, ce, kl, _, means, _ = om.step_losses(cfg_o, params, f(tokens), f(lens), f(classes), f(labels), eps)
    rel = lambda a, b: float((a.float().cpu() - b).abs().max() / b.abs().max())
    msg = []
    for precision in ("bf16p3f", "fp32x3"):
        eng = VAEEngine(VAEConfig(dec_type="lstm"), dev, precision=precision)
        eng.arena.load_state(params)
        out = eng.train_step(t(tokens), t(lens), t(classes), t(labels), eps=eps.to(dev), clip_gradient=1.0)
        torch.cuda.synchronize()
        d = {"ce": rel(out["ce"], ce), "kl": rel(out["kl"], kl), "means": rel(out["means"], means)}
        assert max(d.values()) < 1e-3, (precision, d)
        assert float(eng.arena.g.abs().max()) == 0.0 and torch.isfinite(eng.arena.w).all()
        msg.append("%s ce=%.4f kl=%.2f dev=%.1e" % (precision, float(out["ce"].mean()), float(out["kl"].mean()), max(d.values())))
    print("smoke ok: rasteriser bit-exact; VAE train step vs oracle: " + "; ".join(msg))
