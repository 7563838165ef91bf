"""ORACLE — test infrastructure only (tests/, smoke, bench cpu_baseline may import it; the product never does).

Piano-roll VarAutoEncoder step (`--featurisation roll`): the training path "over piano-roll tensors ... with a
sigmoid-BCE reconstruction loss" that BASELINE.json's north_star names.  PARITY UNPINNED for the model side: the
reference's HEAD keeps only remnants of its piano-roll generation — BinaryCrossEntropy over a [B, T, P] label
(/root/reference/music_style_transfer/VarAutoEncoder/loss.py:27-81, PINNED through oracle.model.bce_loss and the
reference-generated loss goldens), the [len, 120] roll visualiser (utils.py:52-61), the LSTM decoder (model.py:131-203) —
and no roll model, so this file IS the specification of how those remnants are wired:

  * a slice's multi-hot pitch vector takes the place of the token id: the Embedding lookups of Encoder (model.py:86) and
    LSTMDecoder (model.py:176) become bias-free Dense layers over [128 pitches | start flag | 3 zero columns];
  * encoder input  = start row followed by the S slices (the start row is the position whose output feeds latent_proj,
    model.py:97-100, as SOS is for tokens); class embedding, sqrt(D) scale, positional encodings, Transformer encoder,
    latent projection, reparameterisation and KL are oracle.model's (model.py:73-104,287-296, loss.py:4-12);
  * decoder        = LSTMDecoder with teacher forcing (input of step t = slice t-1, start row for t = 0), output Dense to
    128 logits per slice;
  * loss           = BinaryCrossEntropy(from_sigmoid=False, label_smoothing, negative_label_downweighting)(logits,
    binary roll) + kl_weight * KL, summed over the batch for the backward (trainer.py:172-177).
"""
import math

import torch

from . import model as om

N_PITCH, ROLL_IN = 128, 132


def param_shapes(cfg):
    D, Z, C, H = cfg.enc_size, cfg.latent, cfg.num_classes, cfg.dec_size
    s = [("encoder.class2hid.weight", (C, D)), ("encoder.roll_embedding.weight", (D, ROLL_IN))]
    for l in range(cfg.enc_layers):
        s += list(om._tf_layer_shapes("encoder.encoder.layer%d." % l, D, "ln2").items())
    s += [("encoder.latent_proj.weight", (2 * Z, D)), ("encoder.latent_proj.bias", (2 * Z,)),
          ("decoder.latent2hid.weight", (2 * H, Z)), ("decoder.latent2hid.bias", (2 * H,)),
          ("decoder.class2hid.weight", (C, 2 * H)), ("decoder.roll_embedding.weight", (H, ROLL_IN))]
    for l in range(cfg.dec_layers):
        s += [("decoder.decoder.l%d_i2h_weight" % l, (4 * H, H)), ("decoder.decoder.l%d_h2h_weight" % l, (4 * H, H)),
              ("decoder.decoder.l%d_i2h_bias" % l, (4 * H,)), ("decoder.decoder.l%d_h2h_bias" % l, (4 * H,))]
    s += [("decoder.output_layer.weight", (N_PITCH, H)), ("decoder.output_layer.bias", (N_PITCH,))]
    return s


def init_params(cfg, seed=0):
    """Xavier uniform on weights, zeros on biases, ones on gamma (trainer.py:103-105), as oracle.model.init_params."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in param_shapes(cfg):
        if name.endswith("gamma"):
            p[name] = torch.ones(shape)
        elif name.endswith("bias") or name.endswith("beta"):
            p[name] = torch.zeros(shape)
        else:
            scale = math.sqrt(3.0 / ((shape[0] + shape[1]) / 2.0))
            p[name] = (torch.rand(shape, generator=g) * 2.0 - 1.0) * scale
    return p


def roll_features(roll):
    """uint8 / float roll [B, S, 128] -> (Renc [B, S+1, 132], Rdec [B, S, 132]) as msx_roll_features builds them."""
    B, S, _ = roll.shape
    renc = torch.zeros(B, S + 1, ROLL_IN)
    renc[:, 1:, :N_PITCH] = (roll > 0).float()
    renc[:, 0, N_PITCH] = 1.0
    return renc, renc[:, :S, :].clone()


def forward(cfg, p, roll, classes, eps, masks=None):
    """-> (logits [B, S, 128], means, stds)."""
    D = cfg.enc_size
    renc, rdec = roll_features(roll)
    B, T, _ = renc.shape
    x = renc @ p["encoder.roll_embedding.weight"].t() + p["encoder.class2hid.weight"][classes.long()][:, None, :]
    x = torch.sqrt(torch.tensor(float(D))) * x + om.positional_encodings(D, T)
    mask = torch.ones(B, T)
    for l in range(cfg.enc_layers):
        x = om.encoder_layer(x, mask, p, "encoder.encoder.layer%d." % l, cfg.enc_heads, cfg.enc_dropout, masks)
    lat = om.dense(x[:, 0, :], p, "encoder.latent_proj")
    Z = cfg.latent
    means, stds = lat[:, :Z], lat[:, Z:]
    z = means + eps * stds
    h0, c0 = om.lstm_initial_state(cfg, p, z, classes)
    xe = rdec @ p["decoder.roll_embedding.weight"].t()
    for l in range(cfg.dec_layers):
        xe, _, _ = om.lstm_layer(xe, h0, c0, p, "decoder.decoder.l%d_" % l)
        if l + 1 < cfg.dec_layers:
            xe = om.dropout(xe, cfg.dec_dropout, masks, "decoder.decoder.l%d" % l)
    return om.dense(xe, p, "decoder.output_layer"), means, stds


def step_losses(cfg, p, roll, classes, eps, kl_weight=1.0, label_smoothing=0.0, downweight=True, masks=None):
    logits, means, stds = forward(cfg, p, roll, classes, eps, masks)
    bce = om.bce_loss(logits, (roll > 0).float(), False, label_smoothing, downweight)
    kl = om.kl_loss(means, stds)
    return bce + kl_weight * kl, bce, kl, logits, means, stds


def train_step(cfg, p, opt, roll, classes, eps, kl_weight=1.0, label_smoothing=0.0, downweight=True, masks=None):
    for v in p.values():
        v.requires_grad_(True)
        v.grad = None
    loss, bce, kl, logits, means, stds = step_losses(cfg, p, roll, classes, eps, kl_weight, label_smoothing, downweight, masks)
    loss.sum().backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    for v in p.values():
        v.requires_grad_(False)
    opt.step(p, grads, roll.shape[0])
    return loss.detach(), bce.detach(), kl.detach(), logits.detach(), means.detach(), stds.detach(), grads
