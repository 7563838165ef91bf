"""Small end-to-end invocations of every kernel family, sized for compute-sanitizer (scripts/sanitize.sh): one train step per
precision mode and decoder type at B = 6, T = 17, a style-transfer decode and a beam search.  They are ordinary parity
smoke tests too (finite losses, finite parameters)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(B, T):
    g = torch.Generator().manual_seed(B * T)
    tokens = torch.randint(3, 293, (B, T), generator=g)
    tokens[:, 0] = 1
    lens = torch.randint(2, T + 1, (B,), generator=g)
    lens[0] = T
    for b in range(B):
        tokens[b, lens[b]:] = 0
    labels = torch.cat([tokens[:, 1:], torch.zeros(B, 1, dtype=torch.long)], 1)
    classes = torch.randint(0, 2, (B,), generator=g)
    d = lambda t: t.to(torch.int32).cuda().contiguous()
    return d(tokens), d(lens), d(classes), d(labels)


@pytest.mark.parametrize("precision", ["fp32", "fp32x3", "tf32x3f", "bf16x3f", "bf16p3f", "tf32", "bf16"])
@pytest.mark.parametrize("dec_type", ["lstm", "transformer"])
def test_small_train_step(precision, dec_type):
    from musicstyletransfer_b200.engine import VAEConfig, VAEEngine
    # B * T > 128 rows so that the pair-tile tensor GEMMs (and the 3xTF32 kernel) take the layer GEMMs
    eng = VAEEngine(VAEConfig(dec_type=dec_type, enc_dropout=0.1, dec_dropout=0.1), "cuda:0", precision=precision)
    args = _batch(9, 17)
    for _ in range(2):
        out = eng.train_step(*args, global_batch=9, clip_gradient=1.0)
    torch.cuda.synchronize()
    assert torch.isfinite(out["ce"]).all() and torch.isfinite(out["kl"]).all() and torch.isfinite(eng.arena.w).all()


def test_small_inference_paths():
    from musicstyletransfer_b200.engine import VAEConfig, VAEEngine
    eng = VAEEngine(VAEConfig(dec_type="lstm"), "cuda:0", precision="tf32")
    tokens, lens, classes, _ = _batch(5, 9)
    seqs, score = eng.style_transfer(tokens, lens, classes, seed=1)
    bs, bscore = eng.beam_search(tokens, lens, classes, 3)
    torch.cuda.synchronize()
    assert seqs.shape[0] == 5 and bs.shape[0] == 15 and torch.isfinite(score).all() and torch.isfinite(bscore).all()
