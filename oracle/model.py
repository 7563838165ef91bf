"""Oracle: VarAutoEncoder forward / loss / train step / sampling on torch-CPU fp32.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Un-fused, one torch op per reference op.
Reference paths are relative to /root/reference/music_style_transfer/VarAutoEncoder.

MXNet-1.3 operator semantics assumed (the arithmetic lives in the un-vendored dependency
``mxnet-cu90==1.3.0.post0``, requirements.txt:3):
  Dense(x) = x @ W.T + b (weight [units,in_units]);  Embedding = weight[int(idx)];
  LayerNorm over the last axis, biased variance, eps 1e-5;  softmax default axis -1;
  fused LSTM gate order i,f,g,o with i2h/h2h weights [4H,in]/[4H,H] and two bias vectors;
  Dropout(p) in train mode = mask/(1-p);  Xavier() = uniform(+-sqrt(3/((fan_in+fan_out)/2)))
  with fan_in = shape[1]*prod(shape[2:]), fan_out = shape[0]*prod(shape[2:]);
  Adam as in mxnet/optimizer.py (bias correction folded into lr, eps outside the sqrt).
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from .featurise import EOS_ID, NUM_EVENTS, PAD_ID, SOS_ID


class Cfg:
    """Plain config mirror of ModelConfig/EncoderConfig/DecoderConfig/TransformerConfig/LSTMConfig
    (model.py:11-54, transformer.py:8-21)."""

    def __init__(self, vocab=NUM_EVENTS, num_classes=2, enc_size=256, enc_layers=2, enc_heads=8,
                 latent=256, dec_type="lstm", dec_size=128, dec_layers=1, dec_heads=8,
                 enc_dropout=0.0, dec_dropout=0.0):
        self.vocab, self.num_classes = vocab, num_classes
        self.enc_size, self.enc_layers, self.enc_heads = enc_size, enc_layers, enc_heads
        self.latent = latent
        self.dec_type, self.dec_size, self.dec_layers, self.dec_heads = dec_type, dec_size, dec_layers, dec_heads
        self.enc_dropout, self.dec_dropout = enc_dropout, dec_dropout


def toy_cfg():
    """create_toy_model_config, main.py:14-38 (Transformer encoder + Transformer decoder)."""
    return Cfg(vocab=10, num_classes=3, enc_size=32, enc_layers=1, enc_heads=2, latent=16,
               dec_type="transformer", dec_size=32, dec_layers=1, dec_heads=2)


# ---- parameter inventory (names follow the Gluon attribute paths) ----------------------------
def _tf_layer_shapes(prefix, D, ln2):
    s = OrderedDict()
    for n in ("W_k", "W_q", "W_v", "W_proj"):                       # transformer.py:65-68
        s[prefix + "self_attention." + n + ".weight"] = (D, D)
        s[prefix + "self_attention." + n + ".bias"] = (D,)
    s[prefix + "ln1.gamma"] = (D,)                                  # transformer.py:142 / :175
    s[prefix + "ln1.beta"] = (D,)
    s[prefix + "ff.ff1.weight"] = (4 * D, D)                        # transformer.py:36-40, :144-146
    s[prefix + "ff.ff1.bias"] = (4 * D,)
    s[prefix + "ff.ff2.weight"] = (D, 4 * D)
    s[prefix + "ff.ff2.bias"] = (D,)
    s[prefix + ln2 + ".gamma"] = (D,)                               # ln2 (encoder) / ln3 (decoder)
    s[prefix + ln2 + ".beta"] = (D,)
    return s


def param_shapes(cfg):
    s = OrderedDict()
    D, Z, V, C = cfg.enc_size, cfg.latent, cfg.vocab, cfg.num_classes
    s["encoder.class2hid.weight"] = (C, D)                          # model.py:62-63
    s["encoder.encoder_embedding.weight"] = (V, D)                  # model.py:65-66
    for l in range(cfg.enc_layers):
        s.update(_tf_layer_shapes("encoder.encoder.layer%d." % l, D, "ln2"))
    s["encoder.latent_proj.weight"] = (2 * Z, D)                    # model.py:70-71
    s["encoder.latent_proj.bias"] = (2 * Z,)
    H = cfg.dec_size
    if cfg.dec_type == "lstm":                                      # model.py:137-157
        s["decoder.latent2hid.weight"] = (2 * H, Z)
        s["decoder.latent2hid.bias"] = (2 * H,)
        s["decoder.class2hid.weight"] = (C, 2 * H)
        s["decoder.embedding.weight"] = (V, H)
        for l in range(cfg.dec_layers):
            s["decoder.decoder.l%d_i2h_weight" % l] = (4 * H, H)
            s["decoder.decoder.l%d_h2h_weight" % l] = (4 * H, H)
            s["decoder.decoder.l%d_i2h_bias" % l] = (4 * H,)
            s["decoder.decoder.l%d_h2h_bias" % l] = (4 * H,)
    else:                                                           # model.py:212-227
        s["decoder.latent2hid.weight"] = (H, Z)
        s["decoder.latent2hid.bias"] = (H,)
        s["decoder.class2hid.weight"] = (C, H)
        s["decoder.embedding.weight"] = (V, H)
        for l in range(cfg.dec_layers):
            s.update(_tf_layer_shapes("decoder.decoder.layer%d." % l, H, "ln3"))
    s["decoder.output_layer.weight"] = (V, H)
    s["decoder.output_layer.bias"] = (V,)
    return s


def init_params(cfg, seed=0):
    """Trainer._initialize_model, trainer.py:103-105: mx.init.Xavier() (uniform, avg, magnitude 3)
    on every ``*weight`` (incl. embeddings and the fused-RNN weight blocks), zeros for bias/beta,
    ones for gamma (MXNet Initializer name dispatch)."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        if name.endswith("gamma"):
            p[name] = torch.ones(shape)
        elif name.endswith("bias") or name.endswith("beta"):
            p[name] = torch.zeros(shape)
        else:
            fan_out, fan_in = shape[0], shape[1]
            scale = math.sqrt(3.0 / ((fan_in + fan_out) / 2.0))
            p[name] = (torch.rand(shape, generator=g) * 2.0 - 1.0) * scale
    return p


# ---- transformer.py -------------------------------------------------------------------------
def positional_encodings(model_size, max_len):
    """transformer.py:204-211 (float64 table, column index i in the exponent, sin on even columns,
    cos on odd), then cast to float32 (mx.nd.array default dtype)."""
    pe = np.arange(max_len).reshape((-1, 1)) / np.power(
        10000, (2.0 / model_size) * np.arange(model_size).reshape((1, -1)))
    pe[:, 0::2] = np.sin(pe[:, 0::2])
    pe[:, 1::2] = np.cos(pe[:, 1::2])
    return torch.from_numpy(pe.astype(np.float32))


def dense(x, p, name):
    return x @ p[name + ".weight"].t() + p[name + ".bias"]


def layer_norm(x, p, name, eps=1e-5):
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * p[name + ".gamma"] + p[name + ".beta"]


def dropout(x, rate, masks, site):
    """gluon.nn.Dropout in train mode.  ``masks`` (dict site->0/1 tensor) makes parity runs
    deterministic; rate 0 or masks None -> identity."""
    if rate <= 0.0 or masks is None:
        return x
    if masks == "random":      # throughput runs: fresh Bernoulli mask per call, as gluon.nn.Dropout does
        return x * (torch.rand_like(x) >= rate).float() / (1.0 - rate)
    return x * masks[site] / (1.0 - rate)


def attention(x_kv, x_q, kv_mask, p, prefix, num_heads):
    """MultiHeadDotAttention.hybrid_forward, transformer.py:79-126 (no cache).
    S = K Q^T [B,H,T_K,T_Q] (:96), /sqrt(d_h) (:98), + (-1e9 on padded KEY rows, broadcast over q)
    (:106-116), softmax over the LAST axis = the query axis (:100), O = P^T V (:102)."""
    B, T_K, D = x_kv.shape
    T_Q = x_q.shape[1]
    dh = D // num_heads
    K = dense(x_kv, p, prefix + "W_k").reshape(B, T_K, num_heads, dh).transpose(1, 2)
    V = dense(x_kv, p, prefix + "W_v").reshape(B, T_K, num_heads, dh).transpose(1, 2)
    Q = dense(x_q, p, prefix + "W_q").reshape(B, T_Q, num_heads, dh).transpose(1, 2)
    S = K @ Q.transpose(-1, -2)
    S = S / torch.sqrt(torch.tensor(float(dh)))
    mask = torch.where(kv_mask > 0, torch.zeros_like(kv_mask), torch.full_like(kv_mask, -1e9))
    S = S + mask[:, None, :, None]
    P = torch.softmax(S, dim=-1)
    O = P.transpose(-1, -2) @ V
    O = O.transpose(1, 2).reshape(B, T_Q, D)
    return dense(O, p, prefix + "W_proj")


def feed_forward(x, p, prefix, rate, masks, site):
    """DualFeedForward, transformer.py:42-46."""
    h = torch.relu(dense(x, p, prefix + "ff1"))
    h = dropout(h, rate, masks, site + ".ffh")
    return dense(h, p, prefix + "ff2")


def encoder_layer(x, mask, p, prefix, heads, rate, masks):
    """TransformerEncoderLayer.hybrid_forward, transformer.py:151-159 (post-LN)."""
    a = attention(x, x, mask, p, prefix + "self_attention.", heads)
    x = layer_norm(x + dropout(a, rate, masks, prefix + "att"), p, prefix + "ln1")
    f = feed_forward(x, p, prefix + "ff.", rate, masks, prefix)
    return layer_norm(x + dropout(f, rate, masks, prefix + "ff"), p, prefix + "ln2")


def decoder_layer(x_in, mask, p, prefix, heads, rate, masks):
    """TransformerDecoderLayer.hybrid_forward, transformer.py:192-201: self-attention is NOT
    causal (mask_future_timesteps=False :170-174); second residual is ln3(f + drop(f)) (:199-200)."""
    a = attention(x_in, x_in, mask, p, prefix + "self_attention.", heads)
    x = layer_norm(x_in + dropout(a, rate, masks, prefix + "att"), p, prefix + "ln1")
    f = feed_forward(x, p, prefix + "ff.", rate, masks, prefix)
    return layer_norm(f + dropout(f, rate, masks, prefix + "ff"), p, prefix + "ln3")


# ---- model.py -------------------------------------------------------------------------------
def encoder_forward(cfg, p, tokens, classes, masks=None):
    """Encoder.hybrid_forward, model.py:73-104 + TransformerEncoder.hybrid_forward,
    transformer.py:268-273.  tokens/classes are float tensors holding integer ids."""
    D = cfg.enc_size
    mask = (tokens != 0).float()                                               # model.py:81-83
    tok = p["encoder.encoder_embedding.weight"][tokens.long()]                 # :86
    cls = p["encoder.class2hid.weight"][classes.long()]                        # :89
    x = cls[:, None, :] + tok                                                  # :91
    T = tokens.shape[1]
    x = torch.sqrt(torch.tensor(float(D))) * x + positional_encodings(D, T)    # transformer.py:270
    for l in range(cfg.enc_layers):
        x = encoder_layer(x, mask, p, "encoder.encoder.layer%d." % l, cfg.enc_heads,
                          cfg.enc_dropout, masks)
    last = x[:, 0, :]                                                          # model.py:97
    lat = dense(last, p, "encoder.latent_proj")                                # :100
    Z = cfg.latent
    return lat[:, :Z], lat[:, Z:]                                              # :103


def lstm_layer(x, h, c, p, prefix):
    """One layer of gluon.rnn.LSTM (fused RNN op, layout NTC), gates i,f,g,o."""
    Wi, Wh = p[prefix + "i2h_weight"], p[prefix + "h2h_weight"]
    bi, bh = p[prefix + "i2h_bias"], p[prefix + "h2h_bias"]
    H = Wh.shape[1]
    outs = []
    for t in range(x.shape[1]):
        g = x[:, t, :] @ Wi.t() + bi + h @ Wh.t() + bh
        i = torch.sigmoid(g[:, 0 * H:1 * H])
        f = torch.sigmoid(g[:, 1 * H:2 * H])
        gg = torch.tanh(g[:, 2 * H:3 * H])
        o = torch.sigmoid(g[:, 3 * H:4 * H])
        c = f * c + i * gg
        h = o * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1), h, c


def lstm_initial_state(cfg, p, z, classes):
    """LSTMDecoder.get_initial_state, model.py:159-167: same (h0,c0) repeated for every layer."""
    H = cfg.dec_size
    t = dense(z, p, "decoder.latent2hid") + p["decoder.class2hid.weight"][classes.long()]
    return t[:, :H], t[:, H:]


def lstm_decoder_logits(cfg, p, tokens, z, classes, masks=None):
    """LSTMDecoder.forward_train, model.py:172-183 (logits; probs = softmax)."""
    h0, c0 = lstm_initial_state(cfg, p, z, classes)
    x = p["decoder.embedding.weight"][tokens.long()]                           # :176
    for l in range(cfg.dec_layers):
        x, _, _ = lstm_layer(x, h0, c0, p, "decoder.decoder.l%d_" % l)         # :179
        if l + 1 < cfg.dec_layers:
            x = dropout(x, cfg.dec_dropout, masks, "decoder.decoder.l%d" % l)
    return dense(x, p, "decoder.output_layer")                                 # :182


def transformer_decoder_logits(cfg, p, tokens, seq_lens, z, classes, masks=None):
    """Decoder.forward_train, model.py:237-257 + TransformerDecoder.forward_train,
    transformer.py:234-240."""
    Dd = cfg.dec_size
    B, T = tokens.shape
    emb = p["decoder.embedding.weight"][tokens.long()]                         # model.py:241
    s0 = dense(z, p, "decoder.latent2hid") + p["decoder.class2hid.weight"][classes.long()]  # :231
    x = torch.cat([s0[:, None, :], emb], dim=1)                                # :244
    pos = torch.arange(T + 1)[None, :].float()
    mask = (pos < (seq_lens[:, None] + 1)).float()                             # :246-247 SequenceMask
    x = torch.sqrt(torch.tensor(float(Dd))) * x + positional_encodings(Dd, T + 1)  # transformer.py:237
    for l in range(cfg.dec_layers):
        x = decoder_layer(x, mask, p, "decoder.decoder.layer%d." % l, cfg.dec_heads,
                          cfg.dec_dropout, masks)
    x = x[:, 1:, :]                                                            # model.py:253
    return dense(x, p, "decoder.output_layer")                                 # :256


def decoder_logits(cfg, p, tokens, seq_lens, z, classes, masks=None):
    if cfg.dec_type == "lstm":
        return lstm_decoder_logits(cfg, p, tokens, z, classes, masks)
    return transformer_decoder_logits(cfg, p, tokens, seq_lens, z, classes, masks)


def model_forward(cfg, p, tokens, seq_lens, classes, eps, masks=None):
    """Model.hybrid_forward, model.py:287-296.  ``eps`` replaces mx.nd.random_normal (:292)."""
    means, stds = encoder_forward(cfg, p, tokens, classes, masks)
    z = means + eps * stds
    logits = decoder_logits(cfg, p, tokens, seq_lens, z, classes, masks)
    return torch.softmax(logits, dim=-1), means, stds


# ---- loss.py --------------------------------------------------------------------------------
def kl_loss(means, stds):
    """VariationalKLLoss.hybrid_forward, loss.py:8-12."""
    return (0.5 * (stds * stds + means * means - 1 - torch.log(stds * stds))).sum(dim=1)


def softmax_ce(probs, labels):
    """SoftmaxCrossEntropy.hybrid_forward, loss.py:16-23: on probabilities; mean over ALL T
    columns (padding included in the divisor)."""
    mask = (labels != 0).float()
    logp = torch.log(probs)
    picked = -torch.gather(logp, -1, labels.long().unsqueeze(-1)).squeeze(-1)
    return (picked * mask).mean(dim=1)


def bce_loss(pred, label, from_sigmoid=False, label_smoothing=0.0, negative_label_downweighting=True):
    """BinaryCrossEntropy.hybrid_forward, loss.py:38-81 (incl. the (w*bce)*bce quirk :52-54)."""
    if not from_sigmoid:
        pred = torch.sigmoid(pred)
    s_label = (1.0 - label_smoothing) * label + label_smoothing * 0.5
    bce = -1 * (s_label * torch.log(1e-12 + pred) + (1 - s_label) * torch.log(1e-12 + (1.0 - pred)))
    if negative_label_downweighting:
        pos = (label == 1.0).float()
        n_pos = pos.sum(dim=(1, 2))
        n_neg = (1.0 - pos).sum(dim=(1, 2))
        w = (n_pos / (n_neg + 1e-12))[:, None, None]
        bce = torch.where(label == 0.0, (w * bce) * bce, bce)
    return bce.mean(dim=(1, 2))


# ---- trainer.py -----------------------------------------------------------------------------
class Adam:
    """gluon.Trainer('adam', {...}).step(batch_size), trainer.py:94-101,177 ->
    mxnet.optimizer.Adam.update + adam_update kernel: g = clip(g*rescale + wd*w, +-clip);
    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    w -= lr_t * m / (sqrt(v) + eps)."""

    def __init__(self, params, lr=3e-4, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0, clip_gradient=None):
        self.lr, self.b1, self.b2, self.eps, self.wd, self.clip = lr, beta1, beta2, eps, wd, clip_gradient
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, params, grads, batch_size):
        self.t += 1
        coef1 = 1.0 - self.b1 ** self.t
        coef2 = 1.0 - self.b2 ** self.t
        lr_t = self.lr * math.sqrt(coef2) / coef1
        rescale = 1.0 / batch_size
        with torch.no_grad():
            for k, w in params.items():
                g = grads[k] * rescale + self.wd * w
                if self.clip is not None:
                    g = torch.clamp(g, -self.clip, self.clip)
                self.m[k].mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                self.v[k].mul_(self.b2).add_(g * g, alpha=1.0 - self.b2)
                w.sub_(lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps))


def step_losses(cfg, p, tokens, seq_lens, classes, labels, eps, kl_weight=1.0, masks=None):
    """Forward half of Trainer._step, trainer.py:167-172.  Returns (loss[B], ce[B], kl[B], probs, means, stds)."""
    probs, means, stds = model_forward(cfg, p, tokens, seq_lens, classes, eps, masks)
    ce = softmax_ce(probs, labels)
    kl = kl_loss(means, stds)
    return ce + kl_weight * kl, ce, kl, probs, means, stds


def train_step(cfg, p, opt, tokens, seq_lens, classes, labels, eps, kl_weight=1.0, masks=None,
               batch_size=None):
    """Trainer._step, trainer.py:155-179: loss.backward() with head gradient ones (sum over the
    batch), optimizer.step(batch_size)."""
    for v in p.values():
        v.requires_grad_(True)
        v.grad = None
    loss, ce, kl, probs, means, stds = step_losses(cfg, p, tokens, seq_lens, classes, labels, eps,
                                                   kl_weight, masks)
    loss.sum().backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    for v in p.values():
        v.requires_grad_(False)
    opt.step(p, grads, batch_size if batch_size is not None else tokens.shape[0])
    return loss.detach(), ce.detach(), kl.detach(), probs.detach(), means.detach(), stds.detach(), grads


def step_metrics(probs, labels, top_k=5):
    """Trainer._update_metrics, trainer.py:181-186: Perplexity(ignore_label=0) sums -log p[label]
    over non-PAD labels; metrics.Accuracy :49-74 masked arg-max matches; TopKAccuracy.
    Returns (sum_nll, n_tokens, n_correct, n_topk)."""
    mask = labels != 0
    picked = torch.gather(probs, -1, labels.long().unsqueeze(-1)).squeeze(-1)
    nll = -(torch.log(torch.clamp(picked, min=1e-10)) * mask).sum()
    correct = ((probs.argmax(dim=-1) == labels.long()) & mask).sum()
    topk = probs.topk(min(top_k, probs.shape[-1]), dim=-1).indices
    in_topk = ((topk == labels.long().unsqueeze(-1)).any(dim=-1) & mask).sum()
    return float(nll), int(mask.sum()), int(correct), int(in_topk)


# ---- sampler.py (A12) -----------------------------------------------------------------------
def lstm_step(cfg, p, tok, h, c):
    """LSTMDecoder.forward_inference, model.py:185-203 (one step, all layers share nothing: n_layers
    states are carried per layer)."""
    x = p["decoder.embedding.weight"][tok.long()][:, None, :]
    hs, cs = [], []
    for l in range(cfg.dec_layers):
        x, hn, cn = lstm_layer(x, h[l], c[l], p, "decoder.decoder.l%d_" % l)
        hs.append(hn)
        cs.append(cn)
    logits = dense(x[:, 0, :], p, "decoder.output_layer")
    return torch.softmax(logits, dim=-1), hs, cs


def style_transfer_lstm(cfg, p, tokens, classes_target, uniforms):
    """SamplerBase.compute_initial_decoder_state (sampler.py:145-151: classes overwritten BEFORE
    encoding, z = means) + Sampling.sample (sampler.py:161-189) with the LSTM decoder step.
    ``uniforms`` [I_max, B] in [0,1) replace mx.nd.random.multinomial: token = first index whose
    inclusive cumulative probability exceeds u.  Stop test (:186): all rows emitted SOS or PAD."""
    B, T = tokens.shape
    I_max = 2 * T
    means, _ = encoder_forward(cfg, p, tokens, classes_target)
    h0, c0 = lstm_initial_state(cfg, p, means, classes_target)
    h = [h0 for _ in range(cfg.dec_layers)]
    c = [c0 for _ in range(cfg.dec_layers)]
    seq = torch.full((B, 1), float(SOS_ID))
    for i in range(1, I_max):
        probs, h, c = lstm_step(cfg, p, seq[:, -1], h, c)
        cdf = torch.cumsum(probs, dim=-1)
        nxt = (cdf <= uniforms[i][:, None]).sum(dim=-1).clamp(max=probs.shape[-1] - 1).float()
        seq = torch.cat([seq, nxt[:, None]], dim=1)
        if int(((nxt == SOS_ID) | (nxt == PAD_ID)).sum()) == B:
            break
    return seq


def beam_search_lstm(cfg, p, tokens, classes_target, beam_size):
    """BeamSearchSampler.sample (sampler.py:192-257) on the LSTM decoder, evident intent (the loop at HEAD re-takes the
    previous states, adds the kept score twice and clobbers its step index):
    candidate = score[hyp] - log p[hyp, v]; a hypothesis whose last token is EOS or PAD is frozen (one candidate, PAD,
    at its own score); at step 1 only beam 0 is expanded (all beams are copies); the beam_size smallest candidates win,
    ties by the smaller flat index hyp * V + v; states, sequences and scores are reordered by the winners; stop when
    every current token is EOS or PAD (:250).  Returns (sequences [B*beam, <= 2T], scores [B*beam])."""
    B, T = tokens.shape
    K, I_max = beam_size, 2 * T
    means, _ = encoder_forward(cfg, p, tokens, classes_target)
    h0, c0 = lstm_initial_state(cfg, p, means, classes_target)
    h = [h0.repeat_interleave(K, dim=0) for _ in range(cfg.dec_layers)]
    c = [c0.repeat_interleave(K, dim=0) for _ in range(cfg.dec_layers)]
    seq = torch.full((B * K, I_max), float(PAD_ID))
    seq[:, 0] = SOS_ID
    scores = torch.zeros(B * K)
    stop = I_max - 1
    for i in range(1, I_max):
        prev = seq[:, i - 1]
        probs, hn, cn = lstm_step(cfg, p, prev, h, c)
        V = probs.shape[-1]
        logp = torch.log_softmax(torch.log(probs), dim=-1)
        cand = scores[:, None] - logp
        fin = (prev == EOS_ID) | (prev == PAD_ID)
        frozen = torch.full_like(cand, float("inf"))
        frozen[:, PAD_ID] = scores
        cand = torch.where(fin[:, None], frozen, cand)
        if i == 1:
            cand.view(B, K, V)[:, 1:, :] = float("inf")
        flat = cand.view(B, K * V)
        order = torch.argsort(flat, dim=1, stable=True)[:, :K]          # ascending, ties by index
        vals = torch.gather(flat, 1, order)
        hyp = (order // V + torch.arange(B)[:, None] * K).reshape(-1)
        word = (order % V).reshape(-1)
        new_scores = torch.where(torch.isinf(vals.reshape(-1)), scores[hyp], vals.reshape(-1))
        seq = seq[hyp].clone()
        seq[:, i] = word.float()
        scores = new_scores
        h = [x[hyp] for x in hn]
        c = [x[hyp] for x in cn]
        if int(((word == EOS_ID) | (word == PAD_ID)).sum()) == B * K:
            stop = i
            break
    return seq[:, :stop + 1], scores


def style_transfer_transformer(cfg, p, tokens, classes_target, uniforms):
    """Sampling.sample (sampler.py:161-189) over Decoder.forward_inference / TransformerDecoder.forward_inference
    (model.py:259-272, transformer.py:242-249) with the evident intent of the (buggy at HEAD)
    incremental path: step 0 feeds [latent prefix, SOS] through the layers as in training; afterwards keys / values
    of every position stay cached and the single new query position makes the softmax over the query axis
    (transformer.py:100) identically 1, so the attended value is the SUM of all cached values.  Parity unpinned."""
    B, T = tokens.shape
    I_max = 2 * T
    D, H = cfg.dec_size, cfg.dec_heads
    means, _ = encoder_forward(cfg, p, tokens, classes_target)
    s0 = dense(means, p, "decoder.latent2hid") + p["decoder.class2hid.weight"][classes_target.long()]
    emb = p["decoder.embedding.weight"]
    pe = positional_encodings(D, I_max + 2)
    x = torch.cat([s0[:, None, :], emb[torch.full((B, 1), SOS_ID).long()]], dim=1)
    x = math.sqrt(float(D)) * x + pe[:2]
    mask = torch.ones(B, 2)
    vsum = []
    for l in range(cfg.dec_layers):
        prefix = "decoder.decoder.layer%d." % l
        vsum.append(dense(x, p, prefix + "self_attention.W_v").sum(dim=1))
        x = decoder_layer(x, mask, p, prefix, H, 0.0, None)
    seq = torch.full((B, 1), float(SOS_ID))

    def emit(h, i, seq):
        probs = torch.softmax(dense(h, p, "decoder.output_layer"), dim=-1)
        cdf = torch.cumsum(probs, dim=-1)
        nxt = (cdf <= uniforms[i][:, None]).sum(dim=-1).clamp(max=probs.shape[-1] - 1).float()
        return torch.cat([seq, nxt[:, None]], dim=1), nxt

    seq, nxt = emit(x[:, 1, :], 1, seq)
    for i in range(2, I_max):
        if int(((nxt == SOS_ID) | (nxt == PAD_ID)).sum()) == B:
            break
        cur = math.sqrt(float(D)) * emb[nxt.long()] + pe[i]
        for l in range(cfg.dec_layers):
            prefix = "decoder.decoder.layer%d." % l
            vsum[l] = vsum[l] + dense(cur, p, prefix + "self_attention.W_v")
            a = dense(vsum[l], p, prefix + "self_attention.W_proj")
            h1 = layer_norm(cur + a, p, prefix + "ln1")
            f = feed_forward(h1, p, prefix + "ff.", 0.0, None, prefix)
            cur = layer_norm(f + f, p, prefix + "ln3")
        seq, nxt = emit(cur, i, seq)
    return seq
