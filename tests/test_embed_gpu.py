"""Embedding backward (msx_embed_bwd_ex): the scatter kernel (small problems, prefix rows) and the gather kernel (large
problems: one CTA per vocabulary row and position segment) vs a float64 index_add on the host."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(tokens, classes, dout, V, C, D, scale):
    B, T = tokens.shape
    d = dout.double().view(B, T, D) * scale
    d_tok = torch.zeros(V, D, dtype=torch.float64)
    d_tok.index_add_(0, tokens.reshape(-1).long(), d.reshape(-1, D))
    d_cls = None
    if classes is not None:
        d_cls = torch.zeros(C, D, dtype=torch.float64)
        d_cls.index_add_(0, classes.long(), d.sum(1))
    return d_tok, d_cls


@pytest.mark.parametrize("B,T,D,with_cls", [(256, 65, 256, True), (512, 33, 128, False), (40, 65, 256, True), (300, 129, 32, True),
                                           (2048, 65, 256, True)])
def test_embed_bwd_scatter_and_gather(B, T, D, with_cls):
    from musicstyletransfer_b200 import ops
    V, C = 293, 2
    g = torch.Generator().manual_seed(B + T)
    # 4/4-like distribution: a third of the positions on four ids, some PAD, the rest uniform
    tokens = torch.randint(3, V, (B, T), generator=g, dtype=torch.int32)
    hot = torch.rand(B, T, generator=g) < 0.33
    tokens[hot] = torch.randint(260, 264, (int(hot.sum()),), generator=g, dtype=torch.int32)
    tokens[torch.rand(B, T, generator=g) < 0.1] = 0
    classes = torch.randint(0, C, (B,), generator=g, dtype=torch.int32) if with_cls else None
    dout = torch.randn(B * T, D, generator=g)
    scale = 16.0 if with_cls else 1.0
    want_tok, want_cls = _ref(tokens, classes, dout, V, C, D, scale)
    d_tok = torch.zeros(V, D, device="cuda")
    d_cls = torch.zeros(C, D, device="cuda") if with_cls else None
    ops.embed_bwd(tokens.cuda(), classes.cuda() if with_cls else None, dout.cuda(), d_tok, d_cls, None, B, T, D, 0, scale, V)
    torch.cuda.synchronize()
    s = float(want_tok.abs().max())
    assert float((d_tok.double().cpu() - want_tok).abs().max()) < 2e-5 * s
    if with_cls:
        assert float((d_cls.double().cpu() - want_cls).abs().max()) < 2e-5 * float(want_cls.abs().max())
    # accumulation semantics: a second call adds on top
    ops.embed_bwd(tokens.cuda(), classes.cuda() if with_cls else None, dout.cuda(), d_tok, d_cls, None, B, T, D, 0, scale, V)
    torch.cuda.synchronize()
    assert float((d_tok.double().cpu() - 2 * want_tok).abs().max()) < 4e-5 * s
