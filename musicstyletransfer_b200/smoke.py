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
    # one train step of a small VarAutoEncoder (Transformer encoder + LSTM decoder)
    cfg_o = om.Cfg(enc_size=64, enc_layers=1, enc_heads=4, latent=32, dec_type="lstm", dec_size=32)
    params = om.init_params(cfg_o, seed=0)
    params["encoder.latent_proj.bias"][32:] = 3.0
    cfg = VAEConfig(enc_size=64, enc_layers=1, enc_heads=4, latent=32, dec_type="lstm", dec_size=32)
    eng = VAEEngine(cfg, dev)
    eng.arena.load_state(params)
    tokens, lens, classes, labels = synth.token_rows_4_4(8, 16, seed=0)
    eps = torch.randn(8, 32, generator=torch.Generator().manual_seed(0))
    out = eng.train_step(t(tokens), t(lens), t(classes), t(labels), eps=eps.to(dev), clip_gradient=1.0)
    torch.cuda.synchronize()
    f = lambda a: torch.from_numpy(a).float()
    loss, ce, kl, _, means, _ = om.step_losses(cfg_o, params, f(tokens), f(lens), f(classes), f(labels), eps)
    np.testing.assert_allclose(out["ce"].cpu().numpy(), ce.numpy(), rtol=1e-3)
    np.testing.assert_allclose(out["kl"].cpu().numpy(), kl.numpy(), rtol=1e-3)
    np.testing.assert_allclose(out["means"].cpu().numpy(), means.numpy(), rtol=1e-3, atol=1e-4)
    print("smoke ok: rasteriser bit-exact, VAE step ce=%.4f kl=%.4f" % (float(out["ce"].mean()), float(out["kl"].mean())))
