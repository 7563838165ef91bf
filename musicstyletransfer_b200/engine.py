"""VarAutoEncoder step engine: flat parameter arena + hand-scheduled forward / backward / Adam.

Host-side orchestration of the hot path the reference runs as
``Trainer._step`` (VarAutoEncoder/trainer.py:155-179): ``Model.hybrid_forward`` (model.py:287-296),
``SoftmaxCrossEntropy`` + ``VariationalKLLoss`` (loss.py), ``loss.backward()`` and
``gluon.Trainer.step(batch_size)``.  Every device operation is a libmsx kernel (ops.py); torch only
owns the buffers and the stream.  The schedule is static, so a whole step is CUDA-graph capturable.
"""
import contextlib
import math
from collections import OrderedDict

import numpy as np
import torch

from . import ops

NUM_EVENTS = 293     # MIDIUtil/defaults.py:58
N_PITCH, ROLL_IN = 128, 132      # piano-roll width (defaults.py:51-54) and its 16-byte padded GEMM operand width
LSTM_SITE = 0x400    # dropout sites between the layers of a stacked LSTM decoder: LSTM_SITE + lower layer index
SITE_STRIDE = 16     # dropout site ids: layer*SITE_STRIDE + {0: attention out, 1: ff hidden, 2: ff out}


# precision mode -> (single-pass tensor GEMMs, split of the forward GEMMs' operands, 3xTF32 backward GEMMs, tensor-core
#                    attention, tensor-core LSTM recurrence).  Storage is fp32 in every mode except the bf16 operand copies of
#                    "bf16".  Forward split: "x3" = 3xTF32 (msx_gemm_tc_x3, ~2^-21 per product), "b3" = bf16 hi + lo on
#                    kind::f16 (msx_gemm_tc_b3, ~2^-17 per product at half the tensor time), None = single pass.
#   fp32      every product exact fp32 on FFMA kernels (reference arithmetic, the slowest)
#   fp32x3    strict fp32 on the tensor cores: all GEMMs 3xTF32, attention / LSTM exact -> every gradient within 1e-3
#   tf32x3f   compensated FORWARD: encoder GEMMs 3xTF32, attention scores 3xTF32 -> loss / KL / latent means within the
#             north star's 1e-3 with 10x margin; backward GEMMs, attention and the LSTM recurrence single-pass TF32
#   bf16x3f   tf32x3f with the forward encoder GEMMs on the bf16x3 kernel (same forward accuracy class, faster)
#   bf16p3f   tf32x3f with the forward encoder GEMMs on the p3 kernel: both operands arrive as bf16 hi / lo PLANES written by
#             the kernels that produce them (embedding, LayerNorm, attention, FF1 epilogue; weights split once per step), the
#             GEMM walks hi*hi + hi*lo + lo*hi on kind::f16 with no conversion stage (~2^-17 per product); the context and
#             the FF hidden activation exist only as planes, their two weight gradients read the hi plane on kind::f16
#   tf32      every tensor-core product single-pass TF32
#   bf16      tf32 with the Transformer layers' GEMM operands stored as bfloat16
PRECISIONS = {
    "fp32": (False, None, False, False, False),
    "fp32x3": (False, "x3", True, False, False),
    "tf32x3f": (True, "x3", False, True, True),
    "bf16x3f": (True, "b3", False, True, True),
    "bf16p3f": (True, "p3", False, True, True),
    "tf32": (True, None, False, True, True),
    "bf16": (True, None, False, True, True),
}


class VAEConfig:
    """Flat mirror of ModelConfig (model.py:11-54) + TransformerConfig (transformer.py:8-21) + LSTMConfig."""

    def __init__(self, vocab=NUM_EVENTS, num_classes=2, enc_size=256, enc_layers=2, enc_heads=8, latent=256,
                 dec_type="lstm", dec_size=128, dec_layers=1, dec_heads=8, enc_dropout=0.0, dec_dropout=0.0,
                 featurisation="events"):
        assert dec_type in ("lstm", "transformer")
        # "events": token ids (the reference's HEAD); "roll": K1's piano-roll windows with the sigmoid-BCE reconstruction
        # loss (loss.py:27-81) — LSTM decoder only, see forward_roll
        assert featurisation in ("events", "roll")
        assert featurisation == "events" or dec_type == "lstm", "the piano-roll path uses the LSTM decoder"
        self.featurisation = featurisation
        assert enc_size % enc_heads == 0
        if dec_type == "transformer":
            assert dec_size % dec_heads == 0
        else:
            # 32 / 64 / 128: recurrence kernels with W_h2h resident per CTA pair (128 also on the tensor cores); any other
            # size up to 512 takes the L2-streaming kernels of lstm.cu
            assert dec_layers >= 1 and 1 <= dec_size <= 512, "LSTM decoder: n_layers >= 1, hidden size 1 .. 512"
        self.vocab, self.num_classes = vocab, num_classes
        self.enc_size, self.enc_layers, self.enc_heads = enc_size, enc_layers, enc_heads
        self.latent = latent
        self.dec_type, self.dec_size, self.dec_layers, self.dec_heads = dec_type, dec_size, dec_layers, dec_heads
        self.enc_dropout, self.dec_dropout = enc_dropout, dec_dropout

    def as_dict(self):
        return dict(self.__dict__)


def _tf_layer_entries(prefix, D, ln2):
    """Arena order inside a transformer layer: the K,Q,V projections are adjacent so that one
    [3D, D] GEMM serves transformer.py:88-93."""
    e = []
    for n in ("W_k", "W_q", "W_v"):
        e.append((prefix + "self_attention." + n + ".weight", (D, D)))
    for n in ("W_k", "W_q", "W_v"):
        e.append((prefix + "self_attention." + n + ".bias", (D,)))
    e += [(prefix + "self_attention.W_proj.weight", (D, D)), (prefix + "self_attention.W_proj.bias", (D,)),
          (prefix + "ln1.gamma", (D,)), (prefix + "ln1.beta", (D,)),
          (prefix + "ff.ff1.weight", (4 * D, D)), (prefix + "ff.ff1.bias", (4 * D,)),
          (prefix + "ff.ff2.weight", (D, 4 * D)), (prefix + "ff.ff2.bias", (D,)),
          (prefix + ln2 + ".gamma", (D,)), (prefix + ln2 + ".beta", (D,))]
    return e


def param_entries(cfg):
    """(name, shape) in arena order.  Names follow the Gluon attribute paths of model.py / transformer.py."""
    D, Z, V, C, H = cfg.enc_size, cfg.latent, cfg.vocab, cfg.num_classes, cfg.dec_size
    if getattr(cfg, "featurisation", "events") == "roll":
        # the Embedding lookups become bias-free Dense layers over [128 pitches | start flag | 3 zero columns]
        e = [("encoder.class2hid.weight", (C, D)), ("encoder.roll_embedding.weight", (D, ROLL_IN))]
        for l in range(cfg.enc_layers):
            e += _tf_layer_entries("encoder.encoder.layer%d." % l, D, "ln2")
        e += [("encoder.latent_proj.weight", (2 * Z, D)), ("encoder.latent_proj.bias", (2 * Z,)),
              ("decoder.latent2hid.weight", (2 * H, Z)), ("decoder.latent2hid.bias", (2 * H,)),
              ("decoder.class2hid.weight", (C, 2 * H)), ("decoder.roll_embedding.weight", (H, ROLL_IN))]
        for l in range(cfg.dec_layers):
            e += [("decoder.decoder.l%d_i2h_weight" % l, (4 * H, H)), ("decoder.decoder.l%d_h2h_weight" % l, (4 * H, H)),
                  ("decoder.decoder.l%d_i2h_bias" % l, (4 * H,)), ("decoder.decoder.l%d_h2h_bias" % l, (4 * H,))]
        return e + [("decoder.output_layer.weight", (N_PITCH, H)), ("decoder.output_layer.bias", (N_PITCH,))]
    e = [("encoder.class2hid.weight", (C, D)), ("encoder.encoder_embedding.weight", (V, D))]
    for l in range(cfg.enc_layers):
        e += _tf_layer_entries("encoder.encoder.layer%d." % l, D, "ln2")
    e += [("encoder.latent_proj.weight", (2 * Z, D)), ("encoder.latent_proj.bias", (2 * Z,))]
    if cfg.dec_type == "lstm":
        e += [("decoder.latent2hid.weight", (2 * H, Z)), ("decoder.latent2hid.bias", (2 * H,)),
              ("decoder.class2hid.weight", (C, 2 * H)), ("decoder.embedding.weight", (V, H))]
        for l in range(cfg.dec_layers):
            e += [("decoder.decoder.l%d_i2h_weight" % l, (4 * H, H)), ("decoder.decoder.l%d_h2h_weight" % l, (4 * H, H)),
                  ("decoder.decoder.l%d_i2h_bias" % l, (4 * H,)), ("decoder.decoder.l%d_h2h_bias" % l, (4 * H,))]
    else:
        e += [("decoder.latent2hid.weight", (H, Z)), ("decoder.latent2hid.bias", (H,)),
              ("decoder.class2hid.weight", (C, H)), ("decoder.embedding.weight", (V, H))]
        for l in range(cfg.dec_layers):
            e += _tf_layer_entries("decoder.decoder.layer%d." % l, H, "ln3")
    e += [("decoder.output_layer.weight", (V, H)), ("decoder.output_layer.bias", (V,))]
    return e


def positional_encodings(model_size, max_len):
    """transformer.py:204-211, computed on the host in float64 exactly as the reference does, cast to fp32."""
    pe = np.arange(max_len).reshape((-1, 1)) / np.power(10000, (2.0 / model_size) * np.arange(model_size).reshape((1, -1)))
    pe[:, 0::2] = np.sin(pe[:, 0::2])
    pe[:, 1::2] = np.cos(pe[:, 1::2])
    return pe.astype(np.float32)


class ParamArena:
    """Parameters, gradients and Adam moments as four flat fp32 arenas (every tensor 16-byte aligned)."""

    def __init__(self, cfg, device):
        self.entries = param_entries(cfg)
        self.offsets = OrderedDict()
        off = 0
        for name, shape in self.entries:
            n = int(np.prod(shape))
            self.offsets[name] = (off, n, shape)
            off += (n + 7) // 8 * 8          # 32-byte aligned fp32 tensors = 16-byte aligned bf16 shadows (TMA bases)
        self.numel = off
        self.n_params = sum(n for _, n, _ in self.offsets.values())
        self.w = torch.zeros(off, dtype=torch.float32, device=device)
        self.g = torch.zeros(off, dtype=torch.float32, device=device)
        self.m = torch.zeros(off, dtype=torch.float32, device=device)
        self.v = torch.zeros(off, dtype=torch.float32, device=device)
        self.adam_state = torch.zeros(4, dtype=torch.float32, device=device)   # [t, lr_t, -, -]
        self.w16 = None                      # bf16 shadow of w at the same element offsets (precision="bf16" only)

    def enable_bf16_shadow(self):
        if self.w16 is None:
            self.w16 = torch.zeros(self.numel, dtype=torch.bfloat16, device=self.w.device)

    def refresh_bf16_shadow(self):
        ops.cast_bf16(self.w, self.w16, self.numel)

    def enable_p3_planes(self):
        """bf16 hi / lo planes of w at the same element offsets (w16 = hi, w16lo = lo): the B operands of the p3 GEMMs."""
        if self.w16 is None:
            self.w16 = torch.zeros(self.numel, dtype=torch.bfloat16, device=self.w.device)
        self.w16lo = torch.zeros(self.numel, dtype=torch.bfloat16, device=self.w.device)

    def refresh_p3_planes(self):
        ops.split_planes(self.w, self.w16, self.w16lo, self.numel)

    def view16lo(self, name):
        off, n, shape = self.offsets[name]
        return self.w16lo[off:off + n].view(shape)

    def span16lo(self, first, last):
        a = self.offsets[first][0]
        off, n, _ = self.offsets[last]
        return self.w16lo[a:off + n]

    def view16(self, name):
        off, n, shape = self.offsets[name]
        return self.w16[off:off + n].view(shape)

    def span16(self, first, last):
        a = self.offsets[first][0]
        off, n, _ = self.offsets[last]
        return self.w16[a:off + n]

    def view(self, name, arena=None):
        off, n, shape = self.offsets[name]
        return (self.w if arena is None else arena)[off:off + n].view(shape)

    def grad(self, name):
        return self.view(name, self.g)

    def span(self, first, last, arena=None):
        """Contiguous slab from tensor `first` through tensor `last` (adjacent in the arena, no padding gaps)."""
        a = self.offsets[first][0]
        off, n, _ = self.offsets[last]
        return (self.w if arena is None else arena)[a:off + n]

    def names(self):
        return list(self.offsets)

    def init_xavier(self, seed=0):
        """Trainer._initialize_model, trainer.py:103-105: mx.init.Xavier() = uniform(+-sqrt(3 / ((fan_in+fan_out)/2)))
        on every weight, zeros on bias / beta, ones on gamma.  Host-side generation (one-off)."""
        g = torch.Generator().manual_seed(seed)
        for name, (off, n, shape) in self.offsets.items():
            if name.endswith("gamma"):
                t = torch.ones(shape)
            elif name.endswith("bias") or name.endswith("beta"):
                t = torch.zeros(shape)
            else:
                scale = math.sqrt(3.0 / ((shape[0] + shape[1]) / 2.0))
                t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * scale
            self.view(name).copy_(t)

    def load_state(self, state):
        for name in self.offsets:
            self.view(name).copy_(torch.as_tensor(state[name]).to(self.w.device, torch.float32))

    def state_dict(self):
        return OrderedDict((name, self.view(name).detach().cpu().clone()) for name in self.offsets)


class _Buffers:
    """Activation / gradient workspace for one (B, T) shape."""

    def __init__(self):
        self.t = {}

    def get(self, name, shape, device, dtype=torch.float32):
        key = (name, tuple(shape), dtype)
        buf = self.t.get(key)
        if buf is None:
            buf = torch.empty(shape, dtype=dtype, device=device)
            self.t[key] = buf
        return buf

    def get_zero(self, name, shape, device, dtype=torch.float32):
        """A buffer that is zero-filled ONCE, when it is created: for gradients of which a step only ever writes the same
        few rows (the SOS rows), so the remaining rows stay zero without a per-step memset."""
        key = (name, tuple(shape), dtype)
        buf = self.t.get(key)
        if buf is None:
            buf = torch.zeros(shape, dtype=dtype, device=device)
            self.t[key] = buf
        return buf


class VAEEngine:
    def __init__(self, cfg, device="cuda:0", seed=0, max_len=1024, precision="fp32", sos_rows_only=True):
        """precision: "fp32" = exact FFMA GEMMs (msx_gemm_f32); "tf32" = tcgen05 tensor-core GEMMs with TF32
        operands and fp32 accumulation (msx_gemm_tc); "bf16" = the tf32 path with the Transformer layers' GEMM operands
        (activations, their gradients and a shadow copy of the weights) stored as bfloat16 in HBM and multiplied by
        tcgen05 kind::f16 (msx_gemm_tc_bf16), fp32 accumulation.  Softmax, LayerNorm, residuals, losses, the LSTM
        recurrence, master weights, gradients and Adam are fp32 in every mode."""
        assert precision in PRECISIONS, precision
        self.precision = precision
        self.bf16 = precision == "bf16"
        # What each mode runs (PRECISIONS): single-pass TF32 / bf16 GEMMs (`tensor`), 3xTF32 GEMMs in the forward and / or
        # backward pass (`x3_fwd`, `x3_bwd`: msx_gemm_tc_x3, fp32-equivalent products; shapes the pair tiles do not cover
        # fall back to msx_gemm_f32), tensor-core attention / LSTM recurrence (`attn_tc`, `lstm_tc`) or the exact FFMA ones.
        self.tensor, self.x3_fwd, self.x3_bwd, self.attn_tc, self.lstm_tc = PRECISIONS[precision]
        self.x3_skip = ()                        # diagnostic: substrings of GEMM names that stay single-pass
        # The encoder output is read at position 0 only (model.py:97-100), so in the top encoder layer everything after the
        # attention is computed for the B SOS rows instead of all B*T rows: same losses, same gradients (the other rows'
        # outputs have no consumer and their gradients are exactly zero), ~30 % less GEMM / LayerNorm work per step.
        # sos_rows_only=False computes every position as the reference does (tests compare the two).
        self.sos_rows_only = bool(sos_rows_only)
        self.cfg = cfg
        self.device = torch.device(device)
        self.arena = ParamArena(cfg, self.device)
        self.arena.init_xavier(seed)
        if self.bf16:
            self.arena.enable_bf16_shadow()
        self.p3 = self.x3_fwd == "p3"
        if self.p3:
            self.arena.enable_p3_planes()
        self.pe_enc = torch.from_numpy(positional_encodings(cfg.enc_size, max_len)).to(self.device)
        self.pe_dec = (torch.from_numpy(positional_encodings(cfg.dec_size, max_len)).to(self.device)
                       if cfg.dec_type == "transformer" else None)
        self.max_len = max_len
        self.ldv = (cfg.vocab + 3) // 4 * 4          # padded leading dimension of the logits
        self.metrics = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._bufs = {}
        # Random streams (dropout masks, eps): every kernel keys its generator by base_seed + the DEVICE step counter
        # step_dev, which equals the number of optimiser steps taken (adam_step ticks it on the stream, eagerly and inside
        # captured graphs alike).  One counter for the whole engine: eager steps, every captured graph and resumed runs
        # walk one seed sequence, no two optimisation steps share masks.
        self.base_seed = 0x5EED0000
        self.dropout_seed = self.base_seed
        self.step_count = 0
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.ctx = None
        self._graphs = {}
        self._hmask_ok = {}                   # layer tag -> the FF1 forward wrote the ReLU bit mask
        # Weight-gradient GEMMs have no consumer before the optimiser step.  With few rows (the reference's own batch of 32:
        # every kernel fills a fraction of the GPU and the step is a chain of ~75 dependent launches) they leave the chain:
        # forked onto a side stream behind the kernel that produced dY, joined before Adam (also inside captured graphs).
        self.wgrad_side_rows = 16384          # fork when the reduction has at most this many rows; 0 disables
        # Programmatic dependent launch (msx_set_pdl): the next kernel's grid is scheduled while the current one drains.
        # Measured (B200, graph replay): B = 32 step 0.724 -> 0.707 ms, B = 2048 step 4.40 -> 4.44 ms (the early-resident
        # CTAs of the next kernel cost the full-GPU kernels more than the hand-over saves), so only small steps take it.
        self.pdl_rows = 16384                 # steps with at most this many rows (B * T) launch programmatically; 0 disables
        self.dec_table = True                 # token-event LSTM decoder: first layer's i2h product as a [V, 4H] table
        self.qkv0_table = True                # token-event encoder: first layer's K|Q|V projection from per-step tables
        self._side = None
        self._side_used = False

    # ------------------------------------------------------------------ helpers
    def _buf(self, B, T):
        b = self._bufs.get((B, T))
        if b is None:
            b = _Buffers()
            self._bufs[(B, T)] = b
        return b

    def _W(self, name):
        return self.arena.view(name)

    def _G(self, name):
        return self.arena.grad(name)

    def _dense_fwd(self, x, ldx, M, name_w, name_b, out, ldo, N, K, relu=False, drop_p=0.0, site=0, accumulate=False,
                   w=None, b=None, mask_out=None, decoder=False, gemm_name=None, force_x3=False):
        """mask_out (tensor path only, N % 32 == 0): int32 [M, N/32] bit mask of (out > 0), the ReLU / dropout mask the
        dgrad of the next layer applies; returns True when it was written.
        decoder: a GEMM of the LSTM decoder (i2h, output layer).  It only feeds the reconstruction loss, which single-pass
        TF32 already matches to ~2e-6 (the mean over T x V log-probabilities averages the operand rounding out), so the
        tf32x3f mode does not spend 3xTF32 on it; the strict fp32x3 mode does."""
        w = self._W(name_w) if w is None else w
        b = (self._W(name_b) if name_b else None) if b is None else b
        want = self.x3_fwd if (self.x3_bwd or not decoder) else None
        if force_x3:
            want = "x3"                          # per-step tables (a few hundred rows): always compensated, in every tensor mode
        if want and self.x3_skip and any(k in (gemm_name or name_w or "") for k in self.x3_skip):
            want = None                          # diagnostic: this GEMM single-pass (profiles/micro/diag_x3_sites.py)
        mode = self._gemm_mode(want, x, ldx, w, K, out, ldo, M, N, K)
        if mode:
            use_mask = mask_out is not None and N % 32 == 0
            if mode == "b3":
                ops.gemm_tc_b3(x, ldx, w, K, out, ldo, M, N, K, bias=b, relu=relu, drop_p=drop_p, seed=self.dropout_seed,
                               site=site, accumulate=accumulate, mask_out=mask_out if use_mask else None, ldmask=N // 32)
            else:
                ops.gemm_tc(x, ldx, 0, w, K, 1, out, ldo, M, N, K, bias=b, relu=relu, drop_p=drop_p, seed=self.dropout_seed,
                            site=site, accumulate=accumulate, mask_out=mask_out if use_mask else None, ldmask=N // 32,
                            x3=mode == "x3")
            return use_mask
        ops.gemm(x, ldx, 0, w, K, 1, out, ldo, M, N, K, bias=b, relu=relu, drop_p=drop_p, seed=self.dropout_seed,
                 site=site, accumulate=accumulate)
        return False

    def _set_pdl(self, rows):
        ops.set_pdl(0 < rows <= self.pdl_rows)

    def _wgrad_stream(self, rows, ok=True):
        """Context for a weight-gradient launch: the side stream (ordered after everything enqueued so far) when the problem
        is small, else the current stream.  ok=False: the caller's dY buffer is written again later in this backward pass."""
        if not ok or rows > self.wgrad_side_rows:
            return contextlib.nullcontext()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._side.wait_event(ev)
        self._side_used = True
        return torch.cuda.stream(self._side)

    def _wgrad_join(self):
        """The optimiser (current stream) waits for the forked weight gradients."""
        if self._side_used:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream().wait_event(ev)
            self._side_used = False

    def _lstm_tc(self, Hd, tv):
        """Tensor-core LSTM recurrence (TF32 mma.sync, W_h2h in registers), H = 128."""
        return self.lstm_tc and ops.lstm_tc_supported(Hd, 2 * Hd, tv, tv[:, Hd:])

    def _gemm_mode(self, want_x3, A, lda, B, ldb, C, ldc, M, N, K):
        """"b3" (bf16x3 on tcgen05, forward form only), "x3" (3xTF32 on tcgen05), "tc" (single-pass TF32 on tcgen05) or None
        (exact FFMA kernel) for one GEMM.  want_x3: None / False, True or "x3", "b3"."""
        if want_x3:                             # "p3": GEMMs outside the planes layers (latent heads, tiny D) run 3xTF32
            if want_x3 == "b3" and ops.gemm_tc_b3_supported(A, lda, B, ldb, C, ldc, M, N, K):
                return "b3"
            if ops.gemm_tc_x3_supported(A, lda, B, ldb, C, ldc, M, N, K):
                return "x3"
            return None                         # narrow outputs of a compensated pass stay exact
        if self.tensor and ops.gemm_tc_supported(A, lda, B, ldb, C, ldc, M, N, K):
            return "tc"
        return None

    def _dense_bwd(self, dy, lddy, M, x, ldx, w, gw, gb, N, K, dx=None, lddx=0, aux=None, ldaux=0, aux_scale=1.0,
                   accumulate_dx=False, dx_colsum=None, skip_wgrad=False, fork_ok=True):
        """dy [M,N] -> gw [N,K] += dy^T x, gb [N] += colsum(dy) (gb=None: the kernel that produced dy already added it),
        dx [M,K] (=|+=) dy w (optionally masked by aux).  dx_colsum: bias-gradient buffer of the layer whose
        pre-activation gradient dx is; returns True when the dgrad epilogue accumulated it (tensor path)."""
        sk = max(ops.wgrad_splitk(N, K, M, self.sms), 2)
        mode = None if skip_wgrad else self._gemm_mode(self.x3_bwd, dy, lddy, x, ldx, gw, K, N, K, M)
        if skip_wgrad:                          # the caller computes the weight gradient (planes layers: bf16 hi plane of x)
            assert gb is None
        else:
            with self._wgrad_stream(M, fork_ok):
                if mode:
                    ops.gemm_tc(dy, lddy, 1, x, ldx, 0, gw, K, N, K, M, splitk=sk, x3=mode == "x3")
                    if gb is not None:
                        ops.colsum(dy, lddy, M, N, gb)
                else:
                    ops.gemm(dy, lddy, 1, x, ldx, 0, gw, K, N, K, M, splitk=sk, colsum=gb)
        fused = False
        if dx is not None:
            mode = self._gemm_mode(self.x3_bwd, dy, lddy, w, K, dx, lddx, M, K, N)
            if mode:
                fused = dx_colsum is not None and not accumulate_dx
                ops.gemm_tc(dy, lddy, 0, w, K, 0, dx, lddx, M, K, N, aux=aux, ldaux=ldaux, aux_scale=aux_scale,
                            accumulate=accumulate_dx, out_colsum=dx_colsum if fused else None, x3=mode == "x3")
            else:
                ops.gemm(dy, lddy, 0, w, K, 0, dx, lddx, M, K, N, aux=aux, ldaux=ldaux, aux_scale=aux_scale,
                         accumulate=accumulate_dx)
        return fused

    # ------------------------------------------------------------------ transformer layer
    def _tf_layer_fwd(self, bf, tag, prefix, x_in, mask, B, T, D, H, p, site0, decoder, sos_only=False, qkv_src=None):
        """One post-LN Transformer layer (transformer.py:151-159 / :192-201).  sos_only: the caller reads the layer's output
        at position 0 of every sequence only (the encoder's top layer: model.py:97-100 takes `out[:, 0, :]`), so everything
        after the attention — projection, LayerNorms, feed-forward — runs on those B rows instead of B*T; attention itself
        still sees every position (keys / values of all rows, and the reference's softmax normalises over the queries).
        Returns [B*T, D], or [B, D] when sos_only."""
        M = B * T
        dev = self.device
        a = self.arena
        qkv = bf.get(tag + "qkv", (M, 3 * D), dev)
        wqkv = a.span(prefix + "self_attention.W_k.weight", prefix + "self_attention.W_v.weight")
        bqkv = a.span(prefix + "self_attention.W_k.bias", prefix + "self_attention.W_v.bias")
        if qkv_src is not None:
            self._qkv0_from_tables(bf, qkv, qkv_src, prefix, B, T, D)
        else:
            self._dense_fwd(x_in, D, M, None, None, qkv, 3 * D, 3 * D, D, w=wqkv, b=bqkv, gemm_name=prefix + "qkv")
        ctx = bf.get(tag + "ctx", (M, D), dev)
        if self.attn_tc and ops.attention_tc_supported(qkv, T, D // H):
            ops.attention_tc_fwd(qkv, mask, ctx, B, T, H, D // H, x3_scores=bool(self.x3_fwd),
                                 q0_only=sos_only and D // H == 32)
        elif self.attn_tc and ops.attention_tcl_supported(qkv, T, D // H):      # 128 < T <= 768: key tiles of 128
            ops.attention_tcl_fwd(qkv, mask, ctx, bf.get(tag + "attn_stats", (B * H * T, 2), dev), B, T, H, D // H,
                                  q0_only=sos_only)
        else:
            ops.attention_fwd(qkv, mask, ctx, B, T, H, D // H)
        if sos_only:
            R, ldr = B, T * D                   # rows b*T of ctx / x_in: read in place through the leading dimension
            xres = bf.get(tag + "xin_c", (B, D), dev)
            ops.rows_strided(x_in, T * D, xres, D, B, D)
        else:
            R, ldr, xres = M, D, x_in
        proj = bf.get(tag + "proj", (R, D), dev)
        self._dense_fwd(ctx, ldr, R, prefix + "self_attention.W_proj.weight", prefix + "self_attention.W_proj.bias",
                        proj, D, D, D)
        x1 = bf.get(tag + "x1", (R, D), dev)
        st1 = bf.get(tag + "st1", (2, R), dev)
        ops.add_ln_fwd(xres, proj, self._W(prefix + "ln1.gamma"), self._W(prefix + "ln1.beta"), x1, st1[0], st1[1], R, D,
                       drop_p=p, seed=self.dropout_seed, site=site0)
        h = bf.get(tag + "h", (R, 4 * D), dev)
        hmask = bf.get(tag + "hmask", (R, (4 * D + 31) // 32), dev, torch.int32)
        self._hmask_ok[tag] = self._dense_fwd(x1, D, R, prefix + "ff.ff1.weight", prefix + "ff.ff1.bias", h, 4 * D, 4 * D, D,
                                              relu=True, drop_p=p, site=site0 + 1, mask_out=hmask)
        f = bf.get(tag + "f", (R, D), dev)
        self._dense_fwd(h, 4 * D, R, prefix + "ff.ff2.weight", prefix + "ff.ff2.bias", f, D, D, 4 * D)
        out = bf.get(tag + "out", (R, D), dev)
        st2 = bf.get(tag + "st2", (2, R), dev)
        ln2 = "ln3" if decoder else "ln2"
        # encoder: ln2(x1 + drop(f)) (transformer.py:158); decoder: ln3(f + drop(f)) (transformer.py:200)
        ops.add_ln_fwd(f if decoder else x1, f, self._W(prefix + ln2 + ".gamma"), self._W(prefix + ln2 + ".beta"), out,
                       st2[0], st2[1], R, D, drop_p=p, seed=self.dropout_seed, site=site0 + 2)
        return out

    def _tf_layer_bwd(self, bf, tag, prefix, x_in, mask, dout, dx_in, B, T, D, H, p, site0, decoder, sos_only=False):
        """dout = grad wrt the layer output ([B*T, D]; [B, D] when sos_only); writes grad wrt x_in into dx_in [B*T, D]."""
        M = B * T
        R = B if sos_only else M
        ldr = T * D if sos_only else D
        dev = self.device
        a = self.arena
        f32, b16 = torch.float32, torch.bfloat16
        # planes layer (_tf_layer_fwd_p3): the context and the hidden activation exist only as bf16 hi / lo planes; their
        # weight gradients read the hi plane together with a bf16 copy of the output gradient (kind::f16, fp32 accumulate)
        p3 = not decoder and self._layer_p3_ok(D)
        qkv, proj = bf.t[(tag + "qkv", (M, 3 * D), f32)], bf.t[(tag + "proj", (R, D), f32)]
        ctx = None if p3 else bf.t[(tag + "ctx", (M, D), f32)]
        x1, st1 = bf.t[(tag + "x1", (R, D), f32)], bf.t[(tag + "st1", (2, R), f32)]
        h, f = None if p3 else bf.t[(tag + "h", (R, 4 * D), f32)], bf.t[(tag + "f", (R, D), f32)]
        st2 = bf.t[(tag + "st2", (2, R), f32)]
        df16 = bf.get(tag + "df16", (R, D), dev, b16) if p3 else None
        dproj16 = bf.get(tag + "dproj16", (R, D), dev, b16) if p3 else None
        xres = bf.t[(tag + "xin_c", (B, D), f32)] if sos_only else x_in
        inv_keep = 1.0 / (1.0 - p) if p > 0 else 1.0
        ln2 = "ln3" if decoder else "ln2"
        dx1 = bf.get(tag + "dx1", (R, D), dev)
        df = bf.get(tag + "df", (R, D), dev)
        if decoder:
            ops.add_ln_bwd(f, f, self._W(prefix + ln2 + ".gamma"), st2[0], st2[1], dout, df, None,
                           self._G(prefix + ln2 + ".gamma"), self._G(prefix + ln2 + ".beta"), R, D, drop_p=p,
                           seed=self.dropout_seed, site=site0 + 2, fuse_xy=True, dybias=self._G(prefix + "ff.ff2.bias"))
        else:
            ops.add_ln_bwd(x1, f, self._W(prefix + ln2 + ".gamma"), st2[0], st2[1], dout, dx1, df if p > 0 else None,
                           self._G(prefix + ln2 + ".gamma"), self._G(prefix + ln2 + ".beta"), R, D, drop_p=p,
                           seed=self.dropout_seed, site=site0 + 2, dybias=self._G(prefix + "ff.ff2.bias"), dy16=df16)
            if p <= 0:
                df = dx1
        # ff2: f = h W2^T + b2
        dh = bf.get(tag + "dh", (R, 4 * D), dev)
        # ReLU / dropout mask of the hidden activation: the bit mask the FF1 forward epilogue wrote (4 B per 32 elements)
        # when it ran on the tensor path, else the activation itself
        if self._hmask_ok.get(tag):
            aux, ldaux = bf.t[(tag + "hmask", (R, (4 * D + 31) // 32), torch.int32)], 4 * D // 32
        else:
            aux, ldaux = h, 4 * D
        if p3:
            self._wgrad16(df16, D, bf.t[(tag + "h_p", (2, R, 4 * D), b16)][0], 4 * D, self._G(prefix + "ff.ff2.weight"), D, 4 * D, R)
        # p == 0: df aliases dx1 (and dproj aliases dres below), which later kernels of this pass accumulate into
        fused = self._dense_bwd(df, D, R, h, 4 * D, self._W(prefix + "ff.ff2.weight"), self._G(prefix + "ff.ff2.weight"),
                                None, D, 4 * D, dx=dh, lddx=4 * D, aux=aux, ldaux=ldaux, aux_scale=inv_keep,
                                dx_colsum=self._G(prefix + "ff.ff1.bias"), skip_wgrad=p3, fork_ok=p > 0)
        # ff1: h = drop(relu(x1 W1^T + b1));  dh already holds d(pre-activation)
        self._dense_bwd(dh, 4 * D, R, x1, D, self._W(prefix + "ff.ff1.weight"), self._G(prefix + "ff.ff1.weight"),
                        None if fused else self._G(prefix + "ff.ff1.bias"), 4 * D, D, dx=dx1, lddx=D,
                        accumulate_dx=not decoder)
        # ln1(x_in + drop(proj)); sos_only: the residual gradient of the B rows is kept aside and added to dx_in at the end
        dres = bf.get(tag + "dxin_c", (B, D), dev) if sos_only else dx_in
        dproj = bf.get(tag + "dproj", (R, D), dev)
        ops.add_ln_bwd(xres, proj, self._W(prefix + "ln1.gamma"), st1[0], st1[1], dx1, dres, dproj if p > 0 else None,
                       self._G(prefix + "ln1.gamma"), self._G(prefix + "ln1.beta"), R, D, drop_p=p,
                       seed=self.dropout_seed, site=site0, dybias=self._G(prefix + "self_attention.W_proj.bias"),
                       dy16=dproj16)
        if p <= 0:
            dproj = dres
        # sos_only: the context gradient is non-zero at rows b*T only; the dgrad writes exactly those rows of a buffer that
        # was zero-filled when it was created and is written by nothing else
        dctx = bf.get_zero(tag + "dctx_sos", (M, D), dev) if sos_only else bf.get(tag + "dctx", (M, D), dev)
        if p3:
            self._wgrad16(dproj16, D, bf.t[(tag + "ctx_p", (2, M, D), b16)][0], ldr,
                          self._G(prefix + "self_attention.W_proj.weight"), D, D, R)
        self._dense_bwd(dproj, D, R, ctx, ldr, self._W(prefix + "self_attention.W_proj.weight"),
                        self._G(prefix + "self_attention.W_proj.weight"), None, D, D, dx=dctx, lddx=ldr, skip_wgrad=p3,
                        fork_ok=p > 0)
        dqkv = bf.get(tag + "dqkv", (M, 3 * D), dev)
        wqkv = a.span(prefix + "self_attention.W_k.weight", prefix + "self_attention.W_v.weight")
        gwqkv = a.span(prefix + "self_attention.W_k.weight", prefix + "self_attention.W_v.weight", a.g)
        gbqkv = a.span(prefix + "self_attention.W_k.bias", prefix + "self_attention.W_v.bias", a.g)
        if self.attn_tc and ops.attention_tc_supported(qkv, T, D // H):
            ops.attention_tc_bwd(qkv, mask, dctx, dqkv, B, T, H, D // H, dbias=gbqkv, q0_only=sos_only)
            gbqkv = None
        elif self.attn_tc and ops.attention_tcl_supported(qkv, T, D // H):
            ops.attention_tcl_bwd(qkv, mask, dctx, bf.t[(tag + "attn_stats", (B * H * T, 2), torch.float32)], dqkv, B, T, H,
                                  D // H, dbias=gbqkv, q0_only=sos_only)
            gbqkv = None
        else:
            ops.attention_bwd(qkv, mask, dctx, dqkv, B, T, H, D // H)
        self._dense_bwd(dqkv, 3 * D, M, x_in, D, wqkv, gwqkv, gbqkv, 3 * D, D, dx=dx_in, lddx=D, accumulate_dx=not sos_only)
        if sos_only:
            ops.rows_strided(dres, D, dx_in, T * D, B, D, add=True)

    # ------------------------------------------------------------------ transformer layer, hi / lo plane operands (p3)
    def _layer_p3_ok(self, D):
        """p3 GEMM operands need K % 64 == 0 and the vectorised LayerNorm path (D % 128 == 0)."""
        return self.p3 and D % 128 == 0

    def _tf_layer_fwd_p3(self, bf, tag, prefix, x_in, x_in_p, mask, B, T, D, H, p, site0, decoder, sos_only=False,
                         qkv_src=None):
        """_tf_layer_fwd with every GEMM operand as bf16 hi / lo planes (msx_gemm_tc_p3): x_in_p [2, B*T, D] comes from the
        kernel that produced x_in; the context and the FF hidden activation exist ONLY as planes; qkv, the projection
        output, f, the residual stream and the LayerNorm arithmetic are fp32.  Returns (out fp32, out planes)."""
        M = B * T
        R = B if sos_only else M
        ldr = T * D if sos_only else D
        dev, a, b16 = self.device, self.arena, torch.bfloat16
        seed = self.dropout_seed
        qkv = bf.get(tag + "qkv", (M, 3 * D), dev)
        kq, vq = prefix + "self_attention.W_k.weight", prefix + "self_attention.W_v.weight"
        bqkv = a.span(prefix + "self_attention.W_k.bias", prefix + "self_attention.W_v.bias")
        if qkv_src is not None:
            self._qkv0_from_tables(bf, qkv, qkv_src, prefix, B, T, D)
        else:
            ops.gemm_tc_p3(x_in_p[0], x_in_p[1], D, a.span16(kq, vq), a.span16lo(kq, vq), D, qkv, 3 * D, M, 3 * D, D, bias=bqkv)
        ctxp = bf.get(tag + "ctx_p", (2, M, D), dev, b16)
        if ops.attention_tc_supported(qkv, T, D // H):
            ops.attention_tc_fwd(qkv, mask, ctxp[0], B, T, H, D // H, x3_scores=True, ctx_lo=ctxp[1],
                                 q0_only=sos_only and D // H == 32)
        elif ops.attention_tcl_supported(qkv, T, D // H):  # 128 < T <= 768: the long-row kernels write the planes themselves
            ops.attention_tcl_fwd(qkv, mask, ctxp[0], bf.get(tag + "attn_stats", (B * H * T, 2), dev), B, T, H, D // H,
                                  q0_only=sos_only, ctx_lo=ctxp[1])
        else:                                           # longer rows: fp32 context from the exact kernel, one split
            ctx = bf.get(tag + "ctx", (M, D), dev)
            ops.attention_fwd(qkv, mask, ctx, B, T, H, D // H)
            ops.split_planes(ctx, ctxp[0], ctxp[1])
        if sos_only:
            xres = bf.get(tag + "xin_c", (B, D), dev)
            ops.rows_strided(x_in, T * D, xres, D, B, D)
        else:
            xres = x_in
        wn = prefix + "self_attention.W_proj.weight"
        proj = bf.get(tag + "proj", (R, D), dev)
        ops.gemm_tc_p3(ctxp[0], ctxp[1], ldr, a.view16(wn), a.view16lo(wn), D, proj, D, R, D, D,
                       bias=self._W(prefix + "self_attention.W_proj.bias"))
        x1 = bf.get(tag + "x1", (R, D), dev)
        x1p = bf.get(tag + "x1_p", (2, R, D), dev, b16)
        st1 = bf.get(tag + "st1", (2, R), dev)
        ops.add_ln_fwd(xres, proj, self._W(prefix + "ln1.gamma"), self._W(prefix + "ln1.beta"), x1, st1[0], st1[1], R, D,
                       drop_p=p, seed=seed, site=site0, out16=x1p[0], out16lo=x1p[1])
        hp = bf.get(tag + "h_p", (2, R, 4 * D), dev, b16)
        hmask = bf.get(tag + "hmask", (R, 4 * D // 32), dev, torch.int32)     # ReLU / dropout bit mask for the FF2 dgrad
        wn = prefix + "ff.ff1.weight"
        ops.gemm_tc_p3(x1p[0], x1p[1], D, a.view16(wn), a.view16lo(wn), D, hp[0], 4 * D, R, 4 * D, D,
                       bias=self._W(prefix + "ff.ff1.bias"), relu=True, drop_p=p, seed=seed, site=site0 + 1, mask_out=hmask,
                       ldmask=4 * D // 32, C_lo=hp[1])
        self._hmask_ok[tag] = True
        wn = prefix + "ff.ff2.weight"
        f = bf.get(tag + "f", (R, D), dev)
        ops.gemm_tc_p3(hp[0], hp[1], 4 * D, a.view16(wn), a.view16lo(wn), 4 * D, f, D, R, D, 4 * D,
                       bias=self._W(prefix + "ff.ff2.bias"))
        out = bf.get(tag + "out", (R, D), dev)
        outp = bf.get(tag + "out_p", (2, R, D), dev, b16)
        st2 = bf.get(tag + "st2", (2, R), dev)
        ln2 = "ln3" if decoder else "ln2"
        ops.add_ln_fwd(f if decoder else x1, f, self._W(prefix + ln2 + ".gamma"), self._W(prefix + ln2 + ".beta"), out,
                       st2[0], st2[1], R, D, drop_p=p, seed=seed, site=site0 + 2, out16=outp[0], out16lo=outp[1])
        return out, outp

    # ------------------------------------------------------------------ transformer layer, bf16 variant
    def _layer16_ok(self, D):
        """bf16 GEMM operands need 16-byte aligned rows (D % 8) and the vectorised LayerNorm path (D % 128)."""
        return self.bf16 and D % 128 == 0

    def _tf_layer_fwd16(self, bf, tag, prefix, x_in, x_in16, mask, B, T, D, H, p, site0, decoder, sos_only=False,
                        qkv_src=None):
        """_tf_layer_fwd with bf16 GEMM operands: x_in16 / ctx / x1 / the FF hidden activation are read by the GEMMs as
        bfloat16 (the hidden activation and the context exist only in bf16), weights come from the bf16 shadow arena;
        qkv, the projection output, f and the LayerNorm arithmetic stay fp32.  sos_only: as in _tf_layer_fwd, everything
        after the attention runs on the B SOS rows.  Returns (out fp32, out bf16)."""
        M = B * T
        R = B if sos_only else M
        ldr = T * D if sos_only else D
        dev, a, b16 = self.device, self.arena, torch.bfloat16
        seed = self.dropout_seed
        qkv = bf.get(tag + "qkv", (M, 3 * D), dev)
        wqkv = a.span16(prefix + "self_attention.W_k.weight", prefix + "self_attention.W_v.weight")
        bqkv = a.span(prefix + "self_attention.W_k.bias", prefix + "self_attention.W_v.bias")
        if qkv_src is not None:
            self._qkv0_from_tables(bf, qkv, qkv_src, prefix, B, T, D)
        else:
            ops.gemm_tc_bf16(x_in16, D, 0, wqkv, D, 1, qkv, 3 * D, M, 3 * D, D, bias=bqkv)
        ctx16 = bf.get(tag + "ctx16", (M, D), dev, b16)
        if ops.attention_tc_supported(qkv, T, D // H):
            ops.attention_tc_fwd(qkv, mask, ctx16, B, T, H, D // H, q0_only=sos_only and D // H == 32)
        elif ops.attention_tcl_supported(qkv, T, D // H):
            ops.attention_tcl_fwd(qkv, mask, ctx16, bf.get(tag + "attn_stats", (B * H * T, 2), dev), B, T, H, D // H,
                                  q0_only=sos_only)
        else:                                           # T > 768: FFMA attention in fp32, then one cast
            ctx = bf.get(tag + "ctx", (M, D), dev)
            ops.attention_fwd(qkv, mask, ctx, B, T, H, D // H)
            ops.cast_bf16(ctx, ctx16)
        if sos_only:
            xres = bf.get(tag + "xin_c", (B, D), dev)
            ops.rows_strided(x_in, T * D, xres, D, B, D)
        else:
            xres = x_in
        # the projection output and f feed only their LayerNorm (forward + backward): bf16 when LayerNorm reads x and y
        # separately (encoder); the decoder's ln3(f + drop(f)) aliases x and y and keeps f fp32
        proj = bf.get(tag + "proj16", (R, D), dev, b16)
        ops.gemm_tc_bf16(ctx16, ldr, 0, a.view16(prefix + "self_attention.W_proj.weight"), D, 1, proj, D, R, D, D,
                         bias=self._W(prefix + "self_attention.W_proj.bias"))
        x1 = bf.get(tag + "x1", (R, D), dev)
        x1h = bf.get(tag + "x1_16", (R, D), dev, b16)
        st1 = bf.get(tag + "st1", (2, R), dev)
        ops.add_ln_fwd(xres, proj, self._W(prefix + "ln1.gamma"), self._W(prefix + "ln1.beta"), x1, st1[0], st1[1], R, D,
                       drop_p=p, seed=seed, site=site0, out16=x1h)
        h16 = bf.get(tag + "h16", (R, 4 * D), dev, b16)
        hmask = bf.get(tag + "hmask", (R, 4 * D // 32), dev, torch.int32)     # ReLU / dropout bit mask for the FF2 dgrad
        ops.gemm_tc_bf16(x1h, D, 0, a.view16(prefix + "ff.ff1.weight"), D, 1, h16, 4 * D, R, 4 * D, D,
                         bias=self._W(prefix + "ff.ff1.bias"), relu=True, drop_p=p, seed=seed, site=site0 + 1,
                         mask_out=hmask, ldmask=4 * D // 32)
        f = bf.get(tag + "f", (R, D), dev) if decoder else bf.get(tag + "f16", (R, D), dev, b16)
        ops.gemm_tc_bf16(h16, 4 * D, 0, a.view16(prefix + "ff.ff2.weight"), 4 * D, 1, f, D, R, D, 4 * D,
                         bias=self._W(prefix + "ff.ff2.bias"))
        out = bf.get(tag + "out", (R, D), dev)
        out16 = bf.get(tag + "out16", (R, D), dev, b16)
        st2 = bf.get(tag + "st2", (2, R), dev)
        ln2 = "ln3" if decoder else "ln2"
        ops.add_ln_fwd(f if decoder else x1, f, self._W(prefix + ln2 + ".gamma"), self._W(prefix + ln2 + ".beta"), out,
                       st2[0], st2[1], R, D, drop_p=p, seed=seed, site=site0 + 2, out16=out16)
        return out, out16

    def _wgrad16(self, dy16, lddy, x16, ldx, gw, N, K, M):
        """gw [N,K] += dy16[M,N]^T x16[M,K] (bf16 operands, fp32 split-K reduce-adds into the gradient arena)."""
        sk = max(ops.wgrad_splitk(N, K, M, self.sms), 2)
        with self._wgrad_stream(M):
            ops.gemm_tc_bf16(dy16, lddy, 1, x16, ldx, 0, gw, K, N, K, M, splitk=sk)

    def _tf_layer_bwd16(self, bf, tag, prefix, x_in, x_in16, mask, dout, dx_in, B, T, D, H, p, site0, decoder, sos_only=False):
        """Backward of _tf_layer_fwd16.  Every gradient that only feeds GEMMs (d f, d hidden, d proj, d qkv) is produced
        directly as bfloat16 by the kernel that computes it (LayerNorm backward, dgrad epilogue, attention backward);
        the residual-stream gradients and all parameter gradients are fp32.  sos_only: dout is [B, D] (see _tf_layer_bwd)."""
        M = B * T
        R = B if sos_only else M
        ldr = T * D if sos_only else D
        dev, a, b16, f32 = self.device, self.arena, torch.bfloat16, torch.float32
        seed = self.dropout_seed
        qkv, proj = bf.t[(tag + "qkv", (M, 3 * D), f32)], bf.t[(tag + "proj16", (R, D), b16)]
        ctx16 = bf.t[(tag + "ctx16", (M, D), b16)]
        x1, x1h = bf.t[(tag + "x1", (R, D), f32)], bf.t[(tag + "x1_16", (R, D), b16)]
        st1, st2 = bf.t[(tag + "st1", (2, R), f32)], bf.t[(tag + "st2", (2, R), f32)]
        h16 = bf.t[(tag + "h16", (R, 4 * D), b16)]
        f = bf.t[(tag + "f", (R, D), f32)] if decoder else bf.t[(tag + "f16", (R, D), b16)]
        xres = bf.t[(tag + "xin_c", (B, D), f32)] if sos_only else x_in
        inv_keep = 1.0 / (1.0 - p) if p > 0 else 1.0
        ln2 = "ln3" if decoder else "ln2"
        dx1 = bf.get(tag + "dx1", (R, D), dev)
        df16 = bf.get(tag + "df16", (R, D), dev, b16)
        if decoder:
            dfull = bf.get(tag + "df", (R, D), dev)
            ops.add_ln_bwd(f, f, self._W(prefix + ln2 + ".gamma"), st2[0], st2[1], dout, dfull, None,
                           self._G(prefix + ln2 + ".gamma"), self._G(prefix + ln2 + ".beta"), R, D, drop_p=p, seed=seed,
                           site=site0 + 2, fuse_xy=True, dybias=self._G(prefix + "ff.ff2.bias"), dy16=df16)
        else:
            ops.add_ln_bwd(x1, f, self._W(prefix + ln2 + ".gamma"), st2[0], st2[1], dout, dx1, None,
                           self._G(prefix + ln2 + ".gamma"), self._G(prefix + ln2 + ".beta"), R, D, drop_p=p, seed=seed,
                           site=site0 + 2, dybias=self._G(prefix + "ff.ff2.bias"), dy16=df16)
        # ff2: f = h W2^T + b2
        self._wgrad16(df16, D, h16, 4 * D, self._G(prefix + "ff.ff2.weight"), D, 4 * D, R)
        dh16 = bf.get(tag + "dh16", (R, 4 * D), dev, b16)
        ops.gemm_tc_bf16(df16, D, 0, a.view16(prefix + "ff.ff2.weight"), 4 * D, 0, dh16, 4 * D, R, 4 * D, D,
                         aux=bf.t[(tag + "hmask", (R, 4 * D // 32), torch.int32)], ldaux=4 * D // 32, aux_scale=inv_keep,
                         out_colsum=self._G(prefix + "ff.ff1.bias"))
        # ff1: h = drop(relu(x1 W1^T + b1)); dh16 holds d(pre-activation)
        self._wgrad16(dh16, 4 * D, x1h, D, self._G(prefix + "ff.ff1.weight"), 4 * D, D, R)
        ops.gemm_tc_bf16(dh16, 4 * D, 0, a.view16(prefix + "ff.ff1.weight"), D, 0, dx1, D, R, D, 4 * D,
                         accumulate=not decoder)
        # ln1(x_in + drop(proj)); sos_only: the residual gradient of the B rows is added to dx_in at the end
        dres = bf.get(tag + "dxin_c", (B, D), dev) if sos_only else dx_in
        dproj16 = bf.get(tag + "dproj16", (R, D), dev, b16)
        ops.add_ln_bwd(xres, proj, self._W(prefix + "ln1.gamma"), st1[0], st1[1], dx1, dres, None,
                       self._G(prefix + "ln1.gamma"), self._G(prefix + "ln1.beta"), R, D, drop_p=p, seed=seed, site=site0,
                       dybias=self._G(prefix + "self_attention.W_proj.bias"), dy16=dproj16)
        self._wgrad16(dproj16, D, ctx16, ldr, self._G(prefix + "self_attention.W_proj.weight"), D, D, R)
        # sos_only: only rows b*T of the context gradient are non-zero; they are written into a buffer zero-filled once
        dctx = bf.get_zero(tag + "dctx_sos", (M, D), dev) if sos_only else bf.get(tag + "dctx", (M, D), dev)
        ops.gemm_tc_bf16(dproj16, D, 0, a.view16(prefix + "self_attention.W_proj.weight"), D, 0, dctx, ldr, R, D, D)
        dqkv16 = bf.get(tag + "dqkv16", (M, 3 * D), dev, b16)
        gbqkv = a.span(prefix + "self_attention.W_k.bias", prefix + "self_attention.W_v.bias", a.g)
        if ops.attention_tc_supported(qkv, T, D // H):
            ops.attention_tc_bwd(qkv, mask, dctx, dqkv16, B, T, H, D // H, dbias=gbqkv, q0_only=sos_only)
        elif ops.attention_tcl_supported(qkv, T, D // H):
            ops.attention_tcl_bwd(qkv, mask, dctx, bf.t[(tag + "attn_stats", (B * H * T, 2), f32)], dqkv16, B, T, H, D // H,
                                  dbias=gbqkv, q0_only=sos_only)
        else:
            dqkv = bf.get(tag + "dqkv", (M, 3 * D), dev)
            ops.attention_bwd(qkv, mask, dctx, dqkv, B, T, H, D // H)
            ops.colsum(dqkv, 3 * D, M, 3 * D, gbqkv)
            ops.cast_bf16(dqkv, dqkv16)
        gwqkv = a.span(prefix + "self_attention.W_k.weight", prefix + "self_attention.W_v.weight", a.g)
        self._wgrad16(dqkv16, 3 * D, x_in16, D, gwqkv, 3 * D, D, M)
        wqkv = a.span16(prefix + "self_attention.W_k.weight", prefix + "self_attention.W_v.weight")
        ops.gemm_tc_bf16(dqkv16, 3 * D, 0, wqkv, D, 0, dx_in, D, M, D, 3 * D, accumulate=not sos_only)
        if sos_only:
            ops.rows_strided(dres, D, dx_in, T * D, B, D, add=True)

    # ------------------------------------------------------------------ encoder
    def _encode(self, bf, tokens, classes, B, T, p_drop):
        """Encoder.hybrid_forward (model.py:73-104) -> (layer inputs/outputs, key mask, lat = [means | stds])."""
        cfg, dev = self.cfg, self.device
        D, Z, V = cfg.enc_size, cfg.latent, cfg.vocab
        M = B * T
        x = bf.get("enc.x0", (M, D), dev)
        mask = bf.get("enc.mask", (M,), dev)
        l16 = self._layer16_ok(D)
        x16 = bf.get("enc.x0_16", (M, D), dev, torch.bfloat16) if l16 else None
        if l16:
            self.arena.refresh_bf16_shadow()          # 8 MB read + 4 MB written per step; always current, also under graphs
        # first layer's K|Q|V projection from tables (msx_rows_from_tables): x0 is a sum of three table rows, so is x0 W^T + b
        a = self.arena
        cname, ename = "encoder.class2hid.weight", "encoder.encoder_embedding.weight"
        C = cfg.num_classes
        qkv_src = None
        if (self.qkv0_table and self.tensor and (3 * D) % 4 == 0 and 3 * D <= 1024 and M >= 8 * (C + V + T) and
                a.offsets[ename][0] == a.offsets[cname][0] + C * D):   # small batches: the GEMM is cheaper than three launches
            qkv_src = (tokens, classes)
        xlo = None
        if self._layer_p3_ok(D) and qkv_src is None:  # hi / lo planes of the embedded rows (the only reader is that projection)
            xp = bf.get("enc.x0_p", (2, M, D), dev, torch.bfloat16)
            x16, xlo = xp[0], xp[1]
        ops.embed_fwd(tokens, classes, None, self._W(ename), self._W(cname), None, self.pe_enc, x, mask, B, T, D, 0,
                      math.sqrt(float(D)), V, out16=x16, out16lo=xlo)
        if xlo is not None:
            x16 = xp
        return self._encode_layers(bf, x, x16, mask, B, T, p_drop, qkv_src=qkv_src)

    def _qkv0_from_tables(self, bf, qkv, qkv_src, prefix, B, T, D):
        """qkv [B*T, 3D] of the first encoder layer = scale * (tab[C + token] + tab[class]) + postab[position]: two GEMMs over
        C + V and T rows (compensated like the projection they replace) and one gather-add pass."""
        cfg, dev, a = self.cfg, self.device, self.arena
        C, V = cfg.num_classes, cfg.vocab
        tokens, classes = qkv_src
        wqkv = a.span(prefix + "self_attention.W_k.weight", prefix + "self_attention.W_v.weight")
        bqkv = a.span(prefix + "self_attention.W_k.bias", prefix + "self_attention.W_v.bias")
        tab = bf.get("enc.qkv0_tab", (C + V, 3 * D), dev)
        postab = bf.get("enc.qkv0_postab", (T, 3 * D), dev)
        src = a.span("encoder.class2hid.weight", "encoder.encoder_embedding.weight")       # [C + V, D], adjacent in the arena
        self._dense_fwd(src, D, C + V, None, None, tab, 3 * D, 3 * D, D, w=wqkv, force_x3=True)
        self._dense_fwd(self.pe_enc, D, T, None, None, postab, 3 * D, 3 * D, D, w=wqkv, b=bqkv, force_x3=True)
        ops.rows_from_tables(tokens, classes, tab, postab, qkv, B, T, 3 * D, C, V, math.sqrt(float(D)))

    def _encode_layers(self, bf, x, x16, mask, B, T, p_drop, qkv_src=None):
        """Transformer encoder layers + latent projection on an embedded input x [B*T, D] (transformer.py:268-273,
        model.py:97-103)."""
        cfg, dev = self.cfg, self.device
        D, Z = cfg.enc_size, cfg.latent
        l16 = x16 is not None and not self._layer_p3_ok(D)
        if self._layer_p3_ok(D):
            self.arena.refresh_p3_planes()               # weights -> hi / lo planes: 8 MB read + 8 MB written per step
        xs = [x]
        self._xs16 = [x16]
        for l in range(cfg.enc_layers):
            qs = qkv_src if l == 0 else None
            if self._layer_p3_ok(D):
                if x16 is None and qs is None:           # an input no kernel wrote as planes (the roll path's dense embedding)
                    x16 = bf.get("enc.x0_p", (2, B * T, D), dev, torch.bfloat16)
                    ops.split_planes(x, x16[0], x16[1])
                x, x16 = self._tf_layer_fwd_p3(bf, "enc%d." % l, "encoder.encoder.layer%d." % l, x, x16, mask, B, T, D,
                                               cfg.enc_heads, p_drop, l * SITE_STRIDE, False, sos_only=self._sos_only(l),
                                               qkv_src=qs)
                self._xs16.append(x16)
            elif l16:
                x, x16 = self._tf_layer_fwd16(bf, "enc%d." % l, "encoder.encoder.layer%d." % l, x, x16, mask, B, T, D,
                                              cfg.enc_heads, p_drop, l * SITE_STRIDE, False, sos_only=self._sos_only(l),
                                              qkv_src=qs)
                self._xs16.append(x16)
            else:
                x = self._tf_layer_fwd(bf, "enc%d." % l, "encoder.encoder.layer%d." % l, x, mask, B, T, D, cfg.enc_heads,
                                       p_drop, l * SITE_STRIDE, False, sos_only=self._sos_only(l), qkv_src=qs)
            xs.append(x)
        lat = bf.get("lat", (B, 2 * Z), dev)
        # latent projection on the SOS position (model.py:97-100): row b of the compact top-layer output, else row b*T
        self._dense_fwd(x, D if self._sos_only(cfg.enc_layers - 1) else T * D, B, "encoder.latent_proj.weight",
                        "encoder.latent_proj.bias", lat, 2 * Z, 2 * Z, D)
        return xs, mask, lat

    def _sos_only(self, l):
        """Encoder layer l is the top layer and runs its row-wise part on the SOS rows only (see _tf_layer_fwd)."""
        return self.sos_rows_only and l == self.cfg.enc_layers - 1

    def decoder_initial_state(self, classes, z):
        """latent2hid(z) + class2hid[classes] (model.py:160 / :231): [B, 2H] for the LSTM decoder, [B, D_d] else."""
        cfg, dev = self.cfg, self.device
        B = z.shape[0]
        n = 2 * cfg.dec_size if cfg.dec_type == "lstm" else cfg.dec_size
        out = torch.empty((B, n), dtype=torch.float32, device=dev)
        ops.embed_fwd(classes, None, None, self._W("decoder.class2hid.weight"), None, None, None, out, None, B, 1, n, 0,
                      1.0, cfg.num_classes)
        self._dense_fwd(z, z.stride(0), B, "decoder.latent2hid.weight", "decoder.latent2hid.bias", out, n, n, cfg.latent,
                        accumulate=True)
        return out

    def encode(self, tokens, classes):
        """Inference-mode encoder: (means [B,Z], stds [B,Z]) as views of one [B,2Z] buffer."""
        B, T = tokens.shape
        self._set_pdl(B * T)
        _, _, lat = self._encode(self._buf(B, T), tokens, classes, B, T, 0.0)
        Z = self.cfg.latent
        return lat[:, :Z], lat[:, Z:]

    # ------------------------------------------------------------------ style transfer / sampling (A12)
    def style_transfer(self, tokens, seq_lens, classes_target, uniforms=None, seed=0):
        """SamplerBase.compute_initial_decoder_state + Sampling.sample (sampler.py:145-151,161-189): the class
        vector is overwritten BEFORE encoding, z = means, then up to 2T-1 autoregressive multinomial steps.
        The all-rows-emitted-SOS/PAD stop test (:186) is evaluated on the host afterwards, which truncates the
        result exactly where the reference's loop would have stopped.  uniforms: optional fp32 [2T, B]."""
        cfg, dev = self.cfg, self.device
        B, T = tokens.shape
        self._set_pdl(B * T)
        Z, V, Hd = cfg.latent, cfg.vocab, cfg.dec_size
        I_max = 2 * T
        bf = self._buf(B, T)
        _, _, lat = self._encode(bf, tokens, classes_target, B, T, 0.0)
        seqs = bf.get("st.seqs", (B, I_max), dev, torch.int32)
        seqs.zero_()
        seqs[:, 0] = 1
        nxt = bf.get("st.next", (B,), dev, torch.int32)
        nxt.fill_(1)
        score = bf.get("st.score", (B,), dev)
        score.zero_()
        logits = bf.get("st.logits", (B, self.ldv), dev)
        W = self._W
        u_at = (lambda i: uniforms[i]) if uniforms is not None else (lambda i: None)
        if cfg.dec_type == "lstm":
            tv = bf.get("st.tvec", (B, 2 * Hd), dev)
            ops.embed_fwd(classes_target, None, None, W("decoder.class2hid.weight"), None, None, None, tv, None, B, 1,
                          2 * Hd, 0, 1.0, cfg.num_classes)
            self._dense_fwd(lat, 2 * Z, B, "decoder.latent2hid.weight", "decoder.latent2hid.bias", tv, 2 * Hd, 2 * Hd, Z,
                            accumulate=True)
            NL = cfg.dec_layers
            hb = [[bf.get("st.h%d_%d" % (l, i), (B, Hd), dev) for i in range(2)] for l in range(NL)]
            cb = [[bf.get("st.c%d_%d" % (l, i), (B, Hd), dev) for i in range(2)] for l in range(NL)]
            hp = bf.get("st.hprev", (B, Hd), dev)
            gates = bf.get("st.gates", (B, 4 * Hd), dev)
            # every layer starts from the same (h0, c0) = the two halves of tv (model.py:159-167)
            state = [(tv, tv[:, Hd:], 2 * Hd) for _ in range(NL)]
            # embedding lookup + i2h Dense of a step (model.py:192-195) = a lookup into the [V, 4H] table
            # emb W_i2h^T + b_i2h, computed once per call: one gather per step instead of a gather and a GEMM
            tab = bf.get("st.i2h_table", (V, 4 * Hd), dev)
            self._dense_fwd(W("decoder.embedding.weight"), Hd, V, "decoder.decoder.l0_i2h_weight",
                            "decoder.decoder.l0_i2h_bias", tab, 4 * Hd, 4 * Hd, Hd, decoder=True)
            for i in range(1, I_max):
                for l in range(NL):
                    if l == 0:
                        ops.embed_fwd(nxt, None, None, tab, None, None, None, gates, None, B, 1, 4 * Hd, 0, 1.0, V)
                    else:                                       # input of layer l = h of the layer below (no dropout at inference)
                        self._dense_fwd(state[l - 1][0], state[l - 1][2], B, "decoder.decoder.l%d_i2h_weight" % l,
                                        "decoder.decoder.l%d_i2h_bias" % l, gates, 4 * Hd, 4 * Hd, Hd, decoder=True)
                    h, c, ld0 = state[l]
                    # one recurrence step; the tensor-core kernel (W_h2h as mma fragments in registers) in the tensor modes
                    step = ops.lstm_tc_fwd if (self.lstm_tc and ops.lstm_tc_supported(Hd, ld0, h, c)) else ops.lstm_fwd
                    step(gates, W("decoder.decoder.l%d_h2h_weight" % l), W("decoder.decoder.l%d_h2h_bias" % l), h, c, ld0,
                         hb[l][i & 1], hp, cb[l][i & 1], B, 1, Hd)                            # model.py:195
                    state[l] = (hb[l][i & 1], cb[l][i & 1], Hd)
                self._dense_fwd(state[NL - 1][0], Hd, B, "decoder.output_layer.weight", "decoder.output_layer.bias", logits,
                                self.ldv, V, Hd, decoder=True)                                # model.py:198
                ops.sample_multinomial(logits, self.ldv, V, u_at(i), seed, i, nxt, score, seqs, I_max, i, B)
        else:
            D = Hd
            H = cfg.dec_heads
            s0 = bf.get("st.s0", (B, D), dev)
            ops.embed_fwd(classes_target, None, None, W("decoder.class2hid.weight"), None, None, None, s0, None, B, 1, D,
                          0, 1.0, cfg.num_classes)
            self._dense_fwd(lat, 2 * Z, B, "decoder.latent2hid.weight", "decoder.latent2hid.bias", s0, D, D, Z,
                            accumulate=True)
            # step 0: positions {latent prefix, SOS} go through the training-path layer (softmax over 2 queries)
            b2 = self._buf(B, -2)
            x = b2.get("st.x0", (B * 2, D), dev)
            m2 = b2.get("st.mask2", (B * 2,), dev)
            two = b2.get("st.len1", (B,), dev, torch.int32)
            two.fill_(1)
            sos = b2.get("st.sos", (B, 1), dev, torch.int32)
            sos.fill_(1)
            ops.embed_fwd(sos, None, two, W("decoder.embedding.weight"), None, s0, self.pe_dec, x, m2, B, 1, D, 1,
                          math.sqrt(float(D)), V)
            vsums = []
            for l in range(cfg.dec_layers):
                prefix = "decoder.decoder.layer%d." % l
                vs = b2.get("st.vsum%d" % l, (B, D), dev)
                wv, bv = W(prefix + "self_attention.W_v.weight"), W(prefix + "self_attention.W_v.bias")
                self._dense_fwd(x, 2 * D, B, None, None, vs, D, D, D, w=wv, b=bv)                     # V of the prefix
                self._dense_fwd(x[1:], 2 * D, B, None, None, vs, D, D, D, w=wv, b=bv, accumulate=True)  # + V of SOS
                vsums.append(vs)
                x = self._tf_layer_fwd(b2, "st%d." % l, prefix, x, m2, B, 2, D, H, 0.0, 0, True)
            self._dense_fwd(x[1:], 2 * D, B, "decoder.output_layer.weight", "decoder.output_layer.bias", logits,
                            self.ldv, V, D)
            ops.sample_multinomial(logits, self.ldv, V, u_at(1), seed, 1, nxt, score, seqs, I_max, 1, B)
            # steps >= 1: one query -> the softmax over the (length-1) query axis is 1 for every key, so the
            # attended value is the running SUM of the cached values (transformer.py:96-103 with a KV cache)
            xi = b2.get("st.xi", (B, D), dev)
            proj = b2.get("st.proj", (B, D), dev)
            x1 = b2.get("st.x1", (B, D), dev)
            hbuf = b2.get("st.h", (B, 4 * D), dev)
            f = b2.get("st.f", (B, D), dev)
            st = b2.get("st.stats", (2, B), dev)
            for i in range(2, I_max):
                ops.embed_fwd(nxt, None, None, W("decoder.embedding.weight"), None, None, self.pe_dec[i:], xi, None, B,
                              1, D, 0, math.sqrt(float(D)), V)
                cur = xi
                for l in range(cfg.dec_layers):
                    prefix = "decoder.decoder.layer%d." % l
                    self._dense_fwd(cur, D, B, prefix + "self_attention.W_v.weight", prefix + "self_attention.W_v.bias",
                                    vsums[l], D, D, D, accumulate=True)
                    self._dense_fwd(vsums[l], D, B, prefix + "self_attention.W_proj.weight",
                                    prefix + "self_attention.W_proj.bias", proj, D, D, D)
                    ops.add_ln_fwd(cur, proj, W(prefix + "ln1.gamma"), W(prefix + "ln1.beta"), x1, st[0], st[1], B, D)
                    self._dense_fwd(x1, D, B, prefix + "ff.ff1.weight", prefix + "ff.ff1.bias", hbuf, 4 * D, 4 * D, D,
                                    relu=True)
                    self._dense_fwd(hbuf, 4 * D, B, prefix + "ff.ff2.weight", prefix + "ff.ff2.bias", f, D, D, 4 * D)
                    out = b2.get("st.out%d" % l, (B, D), dev)
                    ops.add_ln_fwd(f, f, W(prefix + "ln3.gamma"), W(prefix + "ln3.beta"), out, st[0], st[1], B, D)
                    cur = out
                self._dense_fwd(cur, D, B, "decoder.output_layer.weight", "decoder.output_layer.bias", logits, self.ldv,
                                V, D)
                ops.sample_multinomial(logits, self.ldv, V, u_at(i), seed, i, nxt, score, seqs, I_max, i, B)
        host = seqs.cpu()
        done = ((host == 1) | (host == 0)).all(dim=0)            # sampler.py:186
        stop = I_max - 1
        for i in range(1, I_max):
            if bool(done[i]):
                stop = i
                break
        return seqs[:, :stop + 1], score

    def beam_search(self, tokens, seq_lens, classes_target, beam_size):
        """BeamSearchSampler.sample (sampler.py:192-257) on the LSTM decoder, device side: encoder with the target class,
        z = means, beam_size hypotheses per row, up to 2T steps of embed -> i2h GEMM -> LSTM cell -> output GEMM ->
        msx_beam_step (candidate scores, top-k, reordering) -> msx_gather_rows on (h, c).  Rules: csrc/beam.cu (evident
        intent of the reference loop, which is inconsistent at HEAD).  Returns (sequences int32 [B*beam, <= 2T],
        scores fp32 [B*beam]); hypotheses of row b are b*beam .. b*beam+beam-1, best first."""
        cfg, dev = self.cfg, self.device
        assert cfg.dec_type == "lstm", "beam search follows the reference's LSTM-decoder API (sampler.py:222)"
        B, T = tokens.shape
        K, Z, V, Hd, NL = int(beam_size), cfg.latent, cfg.vocab, cfg.dec_size, cfg.dec_layers
        I_max, R = 2 * T, B * int(beam_size)
        bf = self._buf(B, T)
        _, _, lat = self._encode(bf, tokens, classes_target, B, T, 0.0)
        W = self._W
        tv = bf.get("bs.tvec", (B, 2 * Hd), dev)
        ops.embed_fwd(classes_target, None, None, W("decoder.class2hid.weight"), None, None, None, tv, None, B, 1, 2 * Hd, 0,
                      1.0, cfg.num_classes)
        self._dense_fwd(lat, 2 * Z, B, "decoder.latent2hid.weight", "decoder.latent2hid.bias", tv, 2 * Hd, 2 * Hd, Z,
                        accumulate=True)
        bb = self._buf(R, -3)
        # per layer two (h, c) pairs: the states the step reads and the reordered states of the winners
        h = [[bb.get("bs.h%d_%d" % (l, i), (R, Hd), dev) for i in range(2)] for l in range(NL)]
        c = [[bb.get("bs.c%d_%d" % (l, i), (R, Hd), dev) for i in range(2)] for l in range(NL)]
        for l in range(NL):                                            # every layer starts from (h0, c0), model.py:159-167;
            h[l][0].copy_(tv[:, :Hd].repeat_interleave(K, dim=0))      # mx.nd.repeat(state, beam_size, axis=1), sampler.py:212
            c[l][0].copy_(tv[:, Hd:].repeat_interleave(K, dim=0))
        hn = [bb.get("bs.hn%d" % l, (R, Hd), dev) for l in range(NL)]
        cn = [bb.get("bs.cn%d" % l, (R, Hd), dev) for l in range(NL)]
        hp = bb.get("bs.hp", (R, Hd), dev)
        seq = [bb.get("bs.seq%d" % i, (R, I_max), dev, torch.int32) for i in range(2)]
        for sq in seq:
            sq.zero_()
            sq[:, 0] = 1
        score = [bb.get("bs.score%d" % i, (R,), dev) for i in range(2)]
        score[0].zero_()
        nxt = bb.get("bs.next", (R,), dev, torch.int32)
        nxt.fill_(1)
        parent = bb.get("bs.parent", (R,), dev, torch.int32)
        unfinished = bb.get("bs.unfinished", (I_max,), dev, torch.int32)
        unfinished.zero_()
        gates = bb.get("bs.gates", (R, 4 * Hd), dev)
        logits = bb.get("bs.logits", (R, self.ldv), dev)
        cur = 0
        tab = bb.get("bs.i2h_table", (V, 4 * Hd), dev)              # emb W_i2h^T + b_i2h, see style_transfer
        self._dense_fwd(W("decoder.embedding.weight"), Hd, V, "decoder.decoder.l0_i2h_weight", "decoder.decoder.l0_i2h_bias",
                        tab, 4 * Hd, 4 * Hd, Hd)
        for i in range(1, I_max):
            for l in range(NL):
                if l == 0:
                    ops.embed_fwd(nxt, None, None, tab, None, None, None, gates, None, R, 1, 4 * Hd, 0, 1.0, V)
                else:                                               # input of layer l = h of the layer below (model.py:192-195)
                    self._dense_fwd(hn[l - 1], Hd, R, "decoder.decoder.l%d_i2h_weight" % l, "decoder.decoder.l%d_i2h_bias" % l,
                                    gates, 4 * Hd, 4 * Hd, Hd, decoder=True)
                step = ops.lstm_tc_fwd if (self.lstm_tc and ops.lstm_tc_supported(Hd, Hd, h[l][cur], c[l][cur])) else ops.lstm_fwd
                step(gates, W("decoder.decoder.l%d_h2h_weight" % l), W("decoder.decoder.l%d_h2h_bias" % l), h[l][cur], c[l][cur],
                     Hd, hn[l], hp, cn[l], R, 1, Hd)
            self._dense_fwd(hn[NL - 1], Hd, R, "decoder.output_layer.weight", "decoder.output_layer.bias", logits, self.ldv, V, Hd)
            ops.beam_step(logits, self.ldv, V, B, K, seq[cur], seq[cur ^ 1], I_max, i, score[cur], score[cur ^ 1], parent,
                          nxt, unfinished)
            for l in range(NL):
                ops.gather_rows(hn[l], h[l][cur ^ 1], parent, R, Hd)
                ops.gather_rows(cn[l], c[l][cur ^ 1], parent, R, Hd)
            cur ^= 1
        left = unfinished.cpu()[1:]
        done = (left == 0).nonzero()
        stop = int(done[0]) + 1 if done.numel() else I_max - 1      # sampler.py:250: all current tokens EOS / PAD
        return seq[cur][:, :stop + 1].clone(), score[cur].clone()

    # ------------------------------------------------------------------ LSTM decoder (model.py:131-203)
    def _dec_table_ok(self):
        """The token-event LSTM decoder feeds its first layer from the [V, 4H] table emb W_i2h^T + b_i2h (tensor LSTM kernel)."""
        return self.lstm_tc and self.cfg.dec_type == "lstm" and self.cfg.dec_size == 128 and self.dec_table

    def _lstm_decoder_fwd(self, bf, xe, z, classes, B, T, p_drop=0.0, tokens=None):
        """LSTMDecoder.forward_train after the input embedding: initial state latent2hid(z) + class2hid[classes] split into
        (h0, c0) and shared by every layer (model.py:159-167), then per layer the i2h GEMM for all T steps and the persistent
        recurrence; dropout between the layers (gluon.rnn.LSTM(dropout=...), model.py:148-153).  xe [B*T, H] -> hs [B*T, H].
        tokens (xe None): table mode.  The first layer's input is the embedding of `tokens` (model.py:175-179), so its i2h
        product is a row of table [V, 4H] = emb W_i2h^T + b_i2h: one 293-row GEMM per step instead of the embedding gather and
        a [B*T, H] x [H, 4H] GEMM; the recurrence kernel fetches the table rows itself."""
        cfg, dev = self.cfg, self.device
        Z, Hd = cfg.latent, cfg.dec_size
        M = B * T
        tv = bf.get("dec.tvec", (B, 2 * Hd), dev)
        ops.embed_fwd(classes, None, None, self._W("decoder.class2hid.weight"), None, None, None, tv, None, B, 1,
                      2 * Hd, 0, 1.0, cfg.num_classes)
        self._dense_fwd(z, Z, B, "decoder.latent2hid.weight", "decoder.latent2hid.bias", tv, 2 * Hd, 2 * Hd, Z,
                        accumulate=True)
        lstm_fwd = ops.lstm_tc_fwd if self._lstm_tc(Hd, tv) else ops.lstm_fwd
        x = xe
        for l in range(cfg.dec_layers):
            tag = "dec." if l == 0 else "dec.l%d." % l           # layer 0 keeps the single-layer buffer names
            if l > 0:
                xin = bf.get(tag + "xin", (M, Hd), dev)          # dropout(h of the layer below): this layer's input
                if p_drop > 0:
                    ops.dropout(x, xin, p_drop, self.dropout_seed, LSTM_SITE + l - 1)
                    x = xin
                else:
                    x = x                                        # no dropout: the lower layer's hs is read in place
            gates = bf.get(tag + "gates", (M, 4 * Hd), dev)
            hs = bf.get(tag + "hs", (M, Hd), dev)
            hprev = bf.get(tag + "hprev", (M, Hd), dev)
            cs = bf.get(tag + "cs", (M, Hd), dev)
            if l == 0 and x is None:                             # table mode
                V = cfg.vocab
                tab = bf.get("dec.i2h_table", (V, 4 * Hd), dev)
                self._dense_fwd(self._W("decoder.embedding.weight"), Hd, V, "decoder.decoder.l0_i2h_weight",
                                "decoder.decoder.l0_i2h_bias", tab, 4 * Hd, 4 * Hd, Hd, decoder=True)
                ops.lstm_tc_fwd_tab(gates, tokens, tab, self._W("decoder.decoder.l0_h2h_weight"),
                                    self._W("decoder.decoder.l0_h2h_bias"), tv, tv[:, Hd:], 2 * Hd, hs, hprev, cs, B, T, Hd)
            else:
                self._dense_fwd(x, Hd, M, "decoder.decoder.l%d_i2h_weight" % l, "decoder.decoder.l%d_i2h_bias" % l, gates,
                                4 * Hd, 4 * Hd, Hd, decoder=True)
                lstm_fwd(gates, self._W("decoder.decoder.l%d_h2h_weight" % l), self._W("decoder.decoder.l%d_h2h_bias" % l),
                         tv, tv[:, Hd:], 2 * Hd, hs, hprev, cs, B, T, Hd)
            self._lstm_in = getattr(self, "_lstm_in", {})
            self._lstm_in[l] = x
            x = hs
        return x

    def _lstm_decoder_bwd(self, bf, c, ddec, dz, B, T, p_drop=0.0, tokens=None):
        """Backward of _lstm_decoder_fwd: ddec [B*T, H] = gradient of the top layer's hidden states.  Accumulates the LSTM /
        latent2hid / class2hid parameter gradients, writes dz, returns the gradient of the decoder input xe (None in table
        mode, where the embedding / i2h gradients are taken from the per-token sums of d(pre-activations) instead)."""
        cfg, dev = self.cfg, self.device
        Z, Hd = cfg.latent, cfg.dec_size
        M = B * T
        f32 = torch.float32
        tv = bf.t[("dec.tvec", (B, 2 * Hd), f32)]
        dtv = bf.get("dec.dtvec", (B, 2 * Hd), dev)
        lstm_bwd = ops.lstm_tc_bwd if self._lstm_tc(Hd, tv) else ops.lstm_bwd
        dh = ddec
        for l in reversed(range(cfg.dec_layers)):
            tag = "dec." if l == 0 else "dec.l%d." % l
            gates = bf.t[(tag + "gates", (M, 4 * Hd), f32)]
            hprev = bf.t[(tag + "hprev", (M, Hd), f32)]
            cs = bf.t[(tag + "cs", (M, Hd), f32)]
            x_in = self._lstm_in[l]
            # every layer starts from the same (h0, c0): the top layer writes dtv, the others add theirs
            dtv_l = dtv if l == cfg.dec_layers - 1 else bf.get("dec.dtvec_l%d" % l, (B, 2 * Hd), dev)
            lstm_bwd(gates, self._W("decoder.decoder.l%d_h2h_weight" % l), cs, tv[:, Hd:], 2 * Hd, dh, dtv_l, dtv_l[:, Hd:],
                     B, T, Hd, db_i2h=self._G("decoder.decoder.l%d_i2h_bias" % l),
                     db_h2h=self._G("decoder.decoder.l%d_h2h_bias" % l))            # gates now hold d(pre-activations)
            if dtv_l is not dtv:
                ops.rows_strided(dtv_l, 2 * Hd, dtv, 2 * Hd, B, 2 * Hd, add=True)
            self._dense_bwd(gates, 4 * Hd, M, hprev, Hd, None, self._G("decoder.decoder.l%d_h2h_weight" % l), None, 4 * Hd, Hd)
            if l == 0 and x_in is None:
                # table mode: d table[v] = sum of d(pre-activations) over the rows with token v (the sorted-segment row sums
                # of the embedding backward, 512 wide); then d emb += d table W_i2h and d W_i2h += d table^T emb, 293 rows each
                V = cfg.vocab
                dtab = bf.get("dec.d_i2h_table", (V, 4 * Hd), dev)
                dtab.zero_()                                     # 600 KB memset node
                perm, stok = bf.get("dec.tok_perm", (M,), dev, torch.int32), bf.get("dec.tok_sorted", (M,), dev, torch.int32)
                ops.token_sort(tokens, V, perm, stok, bf.get("dec.tok_ws", (3 * V,), dev, torch.int32), M=M)
                ops.rows_sum_by_token(gates, 4 * Hd, 4 * Hd, perm, stok, dtab, M=M)
                self._dense_bwd(dtab, 4 * Hd, V, self._W("decoder.embedding.weight"), Hd,
                                self._W("decoder.decoder.l0_i2h_weight"), self._G("decoder.decoder.l0_i2h_weight"), None,
                                4 * Hd, Hd, dx=self._G("decoder.embedding.weight"), lddx=Hd, accumulate_dx=True)
                dh = None
                continue
            dx = bf.get(tag + "dxe", (M, Hd), dev)
            self._dense_bwd(gates, 4 * Hd, M, x_in, Hd, self._W("decoder.decoder.l%d_i2h_weight" % l),
                            self._G("decoder.decoder.l%d_i2h_weight" % l), None, 4 * Hd, Hd, dx=dx, lddx=Hd)
            if l > 0 and p_drop > 0:
                ops.dropout(dx, dx, p_drop, self.dropout_seed, LSTM_SITE + l - 1)    # the forward's mask, on the gradient
            dh = dx
        ops.embed_bwd(c["classes"], None, dtv, self._G("decoder.class2hid.weight"), None, None, B, 1, 2 * Hd, 0, 1.0,
                      cfg.num_classes)
        self._dense_bwd(dtv, 2 * Hd, B, c["z"], Z, self._W("decoder.latent2hid.weight"),
                        self._G("decoder.latent2hid.weight"), self._G("decoder.latent2hid.bias"), 2 * Hd, Z,
                        dx=dz, lddx=Z)
        return dh

    def _encode_layers_bwd(self, bf, c, dlat, B, T):
        """Backward of _encode_layers: dlat [B, 2Z] -> gradient of the embedded encoder input [B*T, D]."""
        cfg, dev = self.cfg, self.device
        D, Z = cfg.enc_size, cfg.latent
        M = B * T
        xs = c["xs"]
        if self._sos_only(cfg.enc_layers - 1):
            dx = bf.get("enc.dx_top_c", (B, D), dev)             # compact: one row per sequence
            ldt = D
        else:
            # only rows b*T are ever written (by the dgrad below); the rest stays zero from the buffer's creation
            dx = bf.get_zero("enc.dx_top", (M, D), dev)
            ldt = T * D
        self._dense_bwd(dlat, 2 * Z, B, xs[-1], ldt, self._W("encoder.latent_proj.weight"),
                        self._G("encoder.latent_proj.weight"), self._G("encoder.latent_proj.bias"), 2 * Z, D,
                        dx=dx, lddx=ldt)
        for l in reversed(range(cfg.enc_layers)):
            dnext = bf.get("enc%d.dxin" % l, (M, D), dev)
            if self._layer16_ok(D):
                self._tf_layer_bwd16(bf, "enc%d." % l, "encoder.encoder.layer%d." % l, xs[l], c["xs16"][l], c["mask"], dx,
                                     dnext, B, T, D, cfg.enc_heads, c["pe"], l * SITE_STRIDE, False,
                                     sos_only=self._sos_only(l))
            else:
                self._tf_layer_bwd(bf, "enc%d." % l, "encoder.encoder.layer%d." % l, xs[l], c["mask"], dx, dnext, B, T, D,
                                   cfg.enc_heads, c["pe"], l * SITE_STRIDE, False, sos_only=self._sos_only(l))
            dx = dnext
        return dx

    # ------------------------------------------------------------------ forward
    def forward(self, *args, **kwargs):
        """tokens int32 [B,T], seq_lens int32 [B], classes int32 [B], labels int32 [B,T] (optional),
        eps fp32 [B,Z] (None -> Philox N(0,1)).  Returns dict(ce, kl, means, stds[, probs])."""
        try:
            return self._forward(*args, **kwargs)
        finally:
            ops.set_step_counter(None)              # the registration is library-global: never leave it dangling

    def backward(self, *args, **kwargs):
        """Accumulates d(sum_b g_ce[b]*ce_b + g_kl[b]*kl_weight*kl_b)/dparams into the gradient arena
        (trainer.py:172,176: loss = ce + kl_weight*kl, loss.backward() with head gradient ones)."""
        try:
            return self._backward(*args, **kwargs)
        finally:
            ops.set_step_counter(None)

    def _forward(self, tokens, seq_lens, classes, labels=None, eps=None, train=True, want_probs=False,
                 z_override=None, fuse_ce_bwd=False):
        cfg, dev = self.cfg, self.device
        B, T = tokens.shape
        D, Z, V, Hd = cfg.enc_size, cfg.latent, cfg.vocab, cfg.dec_size
        assert T + 1 <= self.max_len
        bf = self._buf(B, T)
        M = B * T
        self._set_pdl(M)
        pe_ = cfg.enc_dropout if train else 0.0
        pd_ = cfg.dec_dropout if train else 0.0
        self.dropout_seed = self.base_seed
        ops.set_step_counter(self.step_dev)         # kernels launched from here on add the device step count to the seed

        xs, mask, lat = self._encode(bf, tokens, classes, B, T, pe_)
        x = xs[-1]
        if eps is None:
            eps = bf.get("eps", (B, Z), dev)
            ops.normal_fill(eps, self.dropout_seed, 0xE95)
        z = bf.get("z", (B, Z), dev)
        kl = bf.get("kl", (B,), dev)
        ops.reparam_kl_fwd(lat, eps, z, kl, B, Z)
        if z_override is not None:
            z = z_override

        # ---- decoder
        if cfg.dec_type == "lstm":
            if self._dec_table_ok():
                hs = self._lstm_decoder_fwd(bf, None, z, classes, B, T, pd_, tokens=tokens)
            else:
                xe = bf.get("dec.xe", (M, Hd), dev)
                ops.embed_fwd(tokens, None, None, self._W("decoder.embedding.weight"), None, None, None, xe, None, B, T, Hd, 0,
                              1.0, V)
                hs = self._lstm_decoder_fwd(bf, xe, z, classes, B, T, pd_)
            dec_out, Td = hs, T
            dmask = None
        else:
            Td = T + 1
            Md = B * Td
            s0 = bf.get("dec.s0", (B, Hd), dev)                # model.py:229-232
            ops.embed_fwd(classes, None, None, self._W("decoder.class2hid.weight"), None, None, None, s0, None, B, 1, Hd,
                          0, 1.0, cfg.num_classes)
            self._dense_fwd(z, Z, B, "decoder.latent2hid.weight", "decoder.latent2hid.bias", s0, Hd, Hd, Z,
                            accumulate=True)
            xd = bf.get("dec.x0", (Md, Hd), dev)
            dmask = bf.get("dec.mask", (Md,), dev)
            d16 = self._layer16_ok(Hd)                  # bf16 variant: the decoder layers take the bf16 operand path too
            xd16 = bf.get("dec.x0_16", (Md, Hd), dev, torch.bfloat16) if d16 else None
            ops.embed_fwd(tokens, None, seq_lens, self._W("decoder.embedding.weight"), None, s0, self.pe_dec, xd, dmask, B,
                          T, Hd, 1, math.sqrt(float(Hd)), V, out16=xd16)
            dxs, dxs16 = [xd], [xd16]
            for l in range(cfg.dec_layers):
                if d16:
                    xd, xd16 = self._tf_layer_fwd16(bf, "dec%d." % l, "decoder.decoder.layer%d." % l, xd, xd16, dmask, B, Td,
                                                    Hd, cfg.dec_heads, pd_, (8 + l) * SITE_STRIDE, True)
                else:
                    xd = self._tf_layer_fwd(bf, "dec%d." % l, "decoder.decoder.layer%d." % l, xd, dmask, B, Td, Hd,
                                            cfg.dec_heads, pd_, (8 + l) * SITE_STRIDE, True)
                dxs.append(xd)
                dxs16.append(xd16)
            dec_out = xd
        Mo = B * Td
        logits = bf.get("logits", (Mo, self.ldv), dev)
        self._dense_fwd(dec_out, Hd, Mo, "decoder.output_layer.weight", "decoder.output_layer.bias", logits, self.ldv, V,
                        Hd, decoder=cfg.dec_type == "lstm")
        out = {"kl": kl, "means": lat[:, :Z], "stds": lat[:, Z:], "z": z}
        lab_full = None
        if labels is not None:
            if cfg.dec_type == "transformer":
                lab_full = bf.get("labels_full", (B, Td), dev, torch.int32)
                ops.prefix_labels(labels, lab_full, B, T)
            else:
                lab_full = labels
            ce = bf.get("ce", (B,), dev)
            lse = bf.get("lse", (Mo,), dev)
            # fuse_ce_bwd (train_step): the cross-entropy backward with head gradient 1 runs in the same pass — the
            # logits are read once and leave this call holding d loss / d logits (backward() then skips msx_ce_bwd)
            ce_fused = bool(fuse_ce_bwd and not want_probs and ops.ce_fwd_bwd_supported(logits, self.ldv, V))
            if ce_fused:
                ops.ce_fwd_bwd(logits, self.ldv, lab_full, ce, lse, self.metrics, B, Td, V, T,
                               dbias=self._G("decoder.output_layer.bias"))
            else:
                ops.ce_fwd(logits, self.ldv, lab_full, ce, lse, self.metrics, B, Td, V, T)
            out["ce"] = ce
        if want_probs:
            probs = torch.empty((Mo, V), dtype=torch.float32, device=dev)
            ops.softmax_rows(logits, self.ldv, probs, Mo, V)
            probs = probs.view(B, Td, V)
            out["probs"] = probs[:, 1:, :] if cfg.dec_type == "transformer" else probs
        self.ctx = dict(B=B, T=T, Td=Td, tokens=tokens, seq_lens=seq_lens, classes=classes, labels=lab_full, eps=eps,
                        xs=xs, mask=mask, lat=lat, z=z, dec_out=dec_out, logits=logits, pe=pe_, pd=pd_, bf=bf,
                        dmask=dmask, dxs=dxs if cfg.dec_type == "transformer" else None, xs16=self._xs16,
                        ce_fused=bool(labels is not None and ce_fused),
                        dxs16=dxs16 if cfg.dec_type == "transformer" else None)
        return out

    # ------------------------------------------------------------------ backward
    def _backward(self, kl_weight=1.0, g_ce=None, g_kl=None):
        c = self.ctx
        cfg, dev, bf = self.cfg, self.device, c["bf"]
        B, T, Td = c["B"], c["T"], c["Td"]
        D, Z, V, Hd = cfg.enc_size, cfg.latent, cfg.vocab, cfg.dec_size
        M, Mo = B * T, B * Td
        ops.set_step_counter(self.step_dev)         # the backward regenerates the forward's dropout masks
        logits = c["logits"]
        lse = bf.t[("lse", (Mo,), torch.float32)]
        fuse_db = V <= 512
        if c.get("ce_fused"):
            assert g_ce is None, "forward(fuse_ce_bwd=True) already wrote d ce / d logits with head gradient 1"
        else:
            ops.ce_bwd(logits, self.ldv, c["labels"], lse, g_ce, B, Td, V, T,       # logits now hold dlogits
                       dbias=self._G("decoder.output_layer.bias") if fuse_db else None)
        ddec = bf.get("ddec", (Mo, Hd), dev)
        self._dense_bwd(logits, self.ldv, Mo, c["dec_out"], Hd, self._W("decoder.output_layer.weight"),
                        self._G("decoder.output_layer.weight"), None if fuse_db else self._G("decoder.output_layer.bias"),
                        V, Hd, dx=ddec, lddx=Hd)
        dz = bf.get("dz", (B, Z), dev)
        if cfg.dec_type == "lstm":
            dxe = self._lstm_decoder_bwd(bf, c, ddec, dz, B, T, c["pd"], tokens=c["tokens"])
            if dxe is not None:
                ops.embed_bwd(c["tokens"], None, dxe, self._G("decoder.embedding.weight"), None, None, B, T, Hd, 0, 1.0, V)
        else:
            dxs = c["dxs"]
            dcur = ddec
            for l in reversed(range(cfg.dec_layers)):
                dnext = bf.get("dec%d.dxin" % l, (Mo, Hd), dev)
                if self._layer16_ok(Hd):
                    self._tf_layer_bwd16(bf, "dec%d." % l, "decoder.decoder.layer%d." % l, dxs[l], c["dxs16"][l], c["dmask"],
                                         dcur, dnext, B, Td, Hd, cfg.dec_heads, c["pd"], (8 + l) * SITE_STRIDE, True)
                else:
                    self._tf_layer_bwd(bf, "dec%d." % l, "decoder.decoder.layer%d." % l, dxs[l], c["dmask"], dcur, dnext, B,
                                       Td, Hd, cfg.dec_heads, c["pd"], (8 + l) * SITE_STRIDE, True)
                dcur = dnext
            ds0 = bf.get("dec.ds0", (B, Hd), dev)
            ops.embed_bwd(c["tokens"], None, dcur, self._G("decoder.embedding.weight"), None, ds0, B, T, Hd, 1,
                          math.sqrt(float(Hd)), V)
            ops.embed_bwd(c["classes"], None, ds0, self._G("decoder.class2hid.weight"), None, None, B, 1, Hd, 0, 1.0,
                          cfg.num_classes)
            self._dense_bwd(ds0, Hd, B, c["z"], Z, self._W("decoder.latent2hid.weight"),
                            self._G("decoder.latent2hid.weight"), self._G("decoder.latent2hid.bias"), Hd, Z, dx=dz, lddx=Z)
        # ---- reparameterisation + KL (model.py:292, loss.py:8-12)
        dlat = bf.get("dlat", (B, 2 * Z), dev)
        ops.reparam_kl_bwd(c["lat"], c["eps"], dz, g_kl, kl_weight, dlat, B, Z)
        dx = self._encode_layers_bwd(bf, c, dlat, B, T)
        ops.embed_bwd(c["tokens"], c["classes"], dx, self._G("encoder.encoder_embedding.weight"),
                      self._G("encoder.class2hid.weight"), None, B, T, D, 0, math.sqrt(float(D)), V)
        self._wgrad_join()

    # ------------------------------------------------------------------ piano-roll step (--featurisation roll)
    def forward_roll(self, roll, classes, eps=None, train=True, label_smoothing=0.0, downweight=True, want_grad=True):
        """The step over K1's piano-roll windows with the sigmoid-BCE reconstruction loss (BASELINE.json north_star;
        BinaryCrossEntropy, loss.py:27-81; model side: derived spec oracle/roll_model.py).  roll uint8 [B, S, 128] (CUDA),
        classes int32 [B].  Encoder input = start row + the S slices (T = S + 1 positions, all real: no key padding), the
        token Embedding becomes a Dense over the multi-hot slice; LSTM decoder with teacher forcing -> logits [B, S, 128].
        Returns dict(bce [B], kl, means, stds, logits); want_grad also leaves d bce / d logits for backward_roll."""
        try:
            return self._forward_roll(roll, classes, eps, train, label_smoothing, downweight, want_grad)
        finally:
            ops.set_step_counter(None)

    def _forward_roll(self, roll, classes, eps, train, label_smoothing, downweight, want_grad):
        cfg, dev = self.cfg, self.device
        assert cfg.featurisation == "roll" and not self.bf16, "forward_roll: engine built for token events / bf16 operands"
        B, S, P = roll.shape
        assert P == N_PITCH and roll.dtype == torch.uint8 and roll.is_contiguous()
        T = S + 1
        D, Z, Hd = cfg.enc_size, cfg.latent, cfg.dec_size
        M, Ms = B * T, B * S
        bf = self._buf(B, T)
        self._set_pdl(M)
        pe_ = cfg.enc_dropout if train else 0.0
        self.dropout_seed = self.base_seed
        ops.set_step_counter(self.step_dev)
        renc = bf.get("roll.renc", (M, ROLL_IN), dev)
        rdec = bf.get("roll.rdec", (Ms, ROLL_IN), dev)
        ops.roll_features(roll, renc, rdec, B, S)
        E = bf.get("roll.E", (M, D), dev)
        self._dense_fwd(renc, ROLL_IN, M, "encoder.roll_embedding.weight", None, E, D, D, ROLL_IN)
        x = bf.get("enc.x0", (M, D), dev)
        ops.embed_dense_fwd(E, classes, self._W("encoder.class2hid.weight"), self.pe_enc, x, B, T, D, math.sqrt(float(D)))
        mask = bf.t.get(("roll.mask", (M,), torch.float32))
        if mask is None:
            mask = bf.get("roll.mask", (M,), dev)
            mask.fill_(1.0)                                     # every slice is a real key (once per shape)
        xs, mask, lat = self._encode_layers(bf, x, None, mask, B, T, pe_)
        if eps is None:
            eps = bf.get("eps", (B, Z), dev)
            ops.normal_fill(eps, self.dropout_seed, 0xE95)
        z = bf.get("z", (B, Z), dev)
        kl = bf.get("kl", (B,), dev)
        ops.reparam_kl_fwd(lat, eps, z, kl, B, Z)
        xe = bf.get("dec.xe", (Ms, Hd), dev)
        self._dense_fwd(rdec, ROLL_IN, Ms, "decoder.roll_embedding.weight", None, xe, Hd, Hd, ROLL_IN)
        pd_ = cfg.dec_dropout if train else 0.0
        hs = self._lstm_decoder_fwd(bf, xe, z, classes, B, S, pd_)
        logits = bf.get("roll.logits", (Ms, N_PITCH), dev)
        self._dense_fwd(hs, Hd, Ms, "decoder.output_layer.weight", "decoder.output_layer.bias", logits, N_PITCH, N_PITCH, Hd)
        bce = bf.get("roll.bce", (B,), dev)
        dlogits = bf.get("roll.dlogits", (Ms, N_PITCH), dev) if want_grad else None
        ops.bce(logits, roll, bce, None, dlogits, B, S * N_PITCH, from_sigmoid=False, label_smoothing=label_smoothing,
                downweight=downweight)
        self.ctx = dict(B=B, T=T, S=S, classes=classes, eps=eps, xs=xs, mask=mask, lat=lat, z=z, hs=hs, pe=pe_, bf=bf,
                        xs16=self._xs16, renc=renc, rdec=rdec, dlogits=dlogits, roll=True, pd=pd_)
        return {"bce": bce, "ce": bce, "kl": kl, "means": lat[:, :Z], "stds": lat[:, Z:], "z": z, "logits": logits.view(B, S, N_PITCH)}

    def backward_roll(self, kl_weight=1.0):
        """d(sum_b bce_b + kl_weight * kl_b) / d params into the gradient arena (trainer.py:172-176)."""
        try:
            c = self.ctx
            cfg, dev, bf = self.cfg, self.device, c["bf"]
            assert c.get("roll") and c["dlogits"] is not None, "backward_roll needs forward_roll(want_grad=True)"
            B, T, S = c["B"], c["T"], c["S"]
            D, Z, Hd = cfg.enc_size, cfg.latent, cfg.dec_size
            M, Ms = B * T, B * S
            ops.set_step_counter(self.step_dev)
            ddec = bf.get("ddec", (Ms, Hd), dev)
            self._dense_bwd(c["dlogits"], N_PITCH, Ms, c["hs"], Hd, self._W("decoder.output_layer.weight"),
                            self._G("decoder.output_layer.weight"), self._G("decoder.output_layer.bias"), N_PITCH, Hd,
                            dx=ddec, lddx=Hd)
            dz = bf.get("dz", (B, Z), dev)
            dxe = self._lstm_decoder_bwd(bf, c, ddec, dz, B, S, c["pd"])
            self._dense_bwd(dxe, Hd, Ms, c["rdec"], ROLL_IN, None, self._G("decoder.roll_embedding.weight"), None, Hd, ROLL_IN)
            dlat = bf.get("dlat", (B, 2 * Z), dev)
            ops.reparam_kl_bwd(c["lat"], c["eps"], dz, None, kl_weight, dlat, B, Z)
            dx = self._encode_layers_bwd(bf, c, dlat, B, T)
            dE = bf.get("roll.dE", (M, D), dev)
            ops.embed_dense_bwd(dx, c["classes"], dE, self._G("encoder.class2hid.weight"), B, T, D, math.sqrt(float(D)))
            self._dense_bwd(dE, D, M, c["renc"], ROLL_IN, None, self._G("encoder.roll_embedding.weight"), None, D, ROLL_IN)
            self._wgrad_join()
        finally:
            ops.set_step_counter(None)

    def train_step_roll(self, roll, classes, eps=None, kl_weight=1.0, global_batch=None, lr=3e-4, clip_gradient=None,
                        label_smoothing=0.0, downweight=True, allreduce=None):
        out = self.forward_roll(roll, classes, eps=eps, train=True, label_smoothing=label_smoothing, downweight=downweight)
        self.backward_roll(kl_weight)
        peer = isinstance(allreduce, str) and allreduce == "peer"
        if allreduce is not None and not peer:
            allreduce(self.arena.g)
        self.adam_step(global_batch or roll.shape[0], lr=lr, clip_gradient=clip_gradient, peer=peer)
        return out

    def train_step_roll_graphed(self, roll, classes, kl_weight=1.0, global_batch=None, lr=3e-4, clip_gradient=None,
                                label_smoothing=0.0, downweight=True):
        """train_step_roll replayed from a CUDA graph (single GPU), one graph per batch shape / hyper-parameter set."""
        B, S, _ = roll.shape
        key = ("roll", B, S, float(kl_weight), global_batch, float(lr), clip_gradient, float(label_smoothing), bool(downweight))
        st = self._graphs.get(key)
        kw = dict(kl_weight=kl_weight, global_batch=global_batch, lr=lr, clip_gradient=clip_gradient,
                  label_smoothing=label_smoothing, downweight=downweight)
        if st is None:
            self._graphs[key] = {"graph": None}
            return self.train_step_roll(roll, classes, **kw)
        if st["graph"] is None:
            from . import lib
            st["in"] = (torch.empty_like(roll), torch.empty_like(classes))
            graph = torch.cuda.CUDAGraph()
            l0 = lib.LAUNCHES
            with torch.cuda.graph(graph):
                out = self.train_step_roll(st["in"][0], st["in"][1], **kw)
            self.step_count -= 1
            st["graph"], st["out"], st["launches"] = graph, out, lib.LAUNCHES - l0
            lib.LAUNCHES = l0
        from . import lib
        st["in"][0].copy_(roll, non_blocking=True)
        st["in"][1].copy_(classes, non_blocking=True)
        st["graph"].replay()
        lib.LAUNCHES += st["launches"]
        self.step_count += 1
        return st["out"]

    # ------------------------------------------------------------------ optimiser
    def adam_step(self, batch_size, lr=3e-4, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0, clip_gradient=None, peer=False):
        """gluon.Trainer.step(batch_size) with MXNet-1.3 Adam (trainer.py:94-101,177); also zeroes the gradients.
        peer=True (after enable_peer_optimizer): the gradients of all ranks are reduced inside the optimiser kernel."""
        a = self.arena
        if peer:
            pr = self._peer
            ops.adam_nvlink_step(a.w, a.g, a.m, a.v, a.numel, a.adam_state, pr["g"], pr["w"], pr["flags"], pr["done"],
                                 pr["rank"], pr["world"], pr["epoch"], lr, beta1, beta2, eps, wd, 1.0 / batch_size,
                                 clip_gradient, zero_grad=True)
        else:
            ops.adam_step(a.w, a.g, a.m, a.v, a.numel, a.adam_state, lr, beta1, beta2, eps, wd, 1.0 / batch_size,
                          clip_gradient, zero_grad=True)
        ops.step_counter_tick(self.step_dev)
        self.step_count += 1

    def set_step_count(self, n):
        """Resume: continue the seed sequence of dropout masks / eps from optimiser step n (checkpoints store it)."""
        self.step_count = int(n)
        self.step_dev.fill_(int(n))

    # ------------------------------------------------------------------ data parallel over NVLink peer memory
    def enable_peer_optimizer(self, group=None):
        """Moves the parameter and gradient arenas into torch symmetric memory (peer-mapped over NVLink) so that
        adam_step(..., peer=True) can run the fused reduce-scatter + Adam + all-gather kernel (msx_adam_nvlink_step)
        instead of ncclAllReduce + a full Adam pass on every rank.  Collective: every rank of `group` must call it,
        before any CUDA graph of the step is captured.  Raises if symmetric memory is unavailable."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        a = self.arena
        n = a.numel
        fl = ops.adam_nvlink_flag_bytes() // 4
        buf = symm_mem.empty(2 * n + fl, dtype=torch.float32, device=self.device)
        hdl = symm_mem.rendezvous(buf, group)
        buf.zero_()
        buf[:n].copy_(a.w)
        a.w, a.g = buf[:n], buf[n:2 * n]
        base = [int(x) for x in hdl.buffer_ptrs]
        self._peer = dict(handle=hdl, buf=buf, rank=int(hdl.rank), world=int(hdl.world_size),
                          w=[b for b in base], g=[b + 4 * n for b in base], flags=[b + 8 * n for b in base],
                          done=torch.zeros(1, dtype=torch.int32, device=self.device),
                          epoch=torch.zeros(1, dtype=torch.int64, device=self.device))
        self._graphs.clear()
        torch.cuda.synchronize(self.device)
        dist.barrier(group)
        return self._peer["world"]

    def train_step_graphed(self, tokens, seq_lens, classes, labels, kl_weight=1.0, global_batch=None, lr=3e-4,
                           clip_gradient=None, allreduce=None):
        """train_step replayed from a CUDA graph (one graph per batch shape and hyper-parameter set).  The schedule of a
        step is static, so the ~66 kernel launches collapse into one graph launch; this is what keeps the reference's own
        B = 32 configuration (scripts/train-vae.sh) from being launch-bound.  The first call with a new key runs eagerly
        (it allocates every buffer), the second captures, later ones only copy the batch into the static input buffers
        and replay.  Dropout masks and eps stay fresh on every replay through the engine-wide device step counter
        (step_dev, msx_set_step_counter) that the captured Adam node ticks.  Returns the same dict as train_step (views of
        static buffers)."""
        B, T = tokens.shape
        peer = isinstance(allreduce, str) and allreduce == "peer"
        key = (B, T, float(kl_weight), global_batch, float(lr), clip_gradient, "peer" if peer else allreduce is not None)
        st = self._graphs.get(key)
        if st is None:
            out = self.train_step(tokens, seq_lens, classes, labels, kl_weight=kl_weight, global_batch=global_batch, lr=lr,
                                  clip_gradient=clip_gradient, allreduce=allreduce)
            self._graphs[key] = {"graph": None}
            return out
        if st["graph"] is None:
            from . import lib
            st["in"] = tuple(torch.empty_like(t) for t in (tokens, seq_lens, classes, labels))
            # with a collective between backward and the optimiser the step is captured as two graphs around an eagerly
            # launched all-reduce (NCCL inside a captured graph ties the graph's lifetime to the communicator's)
            two = allreduce is not None and not peer
            graph, graph2 = torch.cuda.CUDAGraph(), (torch.cuda.CUDAGraph() if two else None)
            l0 = lib.LAUNCHES
            with torch.cuda.graph(graph):
                out = self.forward(*st["in"][:3], st["in"][3], train=True, fuse_ce_bwd=True)
                self.backward(kl_weight)
                if not two:
                    self.adam_step(global_batch or B, lr=lr, clip_gradient=clip_gradient, peer=peer)
            if two:
                with torch.cuda.graph(graph2):
                    self.adam_step(global_batch or B, lr=lr, clip_gradient=clip_gradient)
            self.step_count -= 1                      # the capture pass executed nothing (host mirror of step_dev)
            st["graph"], st["graph2"], st["out"], st["launches"] = graph, graph2, out, lib.LAUNCHES - l0
            lib.LAUNCHES = l0
        from . import lib
        for dst, src in zip(st["in"], (tokens, seq_lens, classes, labels)):
            dst.copy_(src, non_blocking=True)
        st["graph"].replay()
        if st["graph2"] is not None:
            allreduce(self.arena.g)
            st["graph2"].replay()
        lib.LAUNCHES += st["launches"]
        self.step_count += 1
        return st["out"]

    def train_step(self, tokens, seq_lens, classes, labels, eps=None, kl_weight=1.0, global_batch=None, lr=3e-4,
                   clip_gradient=None, allreduce=None):
        """allreduce: None (single GPU), a callable applied to the flat gradient arena (e.g. NCCL all-reduce), or "peer"
        (fused NVLink reduce-scatter + Adam + all-gather, after enable_peer_optimizer)."""
        out = self.forward(tokens, seq_lens, classes, labels, eps=eps, train=True, fuse_ce_bwd=True)
        self.backward(kl_weight)
        peer = isinstance(allreduce, str) and allreduce == "peer"
        if allreduce is not None and not peer:
            allreduce(self.arena.g)
        self.adam_step(global_batch or tokens.shape[0], lr=lr, clip_gradient=clip_gradient, peer=peer)
        return out
