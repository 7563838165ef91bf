"""Model entry points with the reference's names and call signatures (VarAutoEncoder/model.py).

``Model(config)(tokens, seq_lens, classes) -> (probs, means, vars)``; ``Encoder``, ``Decoder`` (Transformer)
and ``LSTMDecoder`` are thin views over one ``VAEEngine`` that owns the flat parameter arena and launches
the kernels.  ``DecoderConfig`` accepts either ``transformer_config`` (HEAD, model.py:22-32) or
``lstm_config`` (what main.py:109-117 passes), which resolves the TypeError that keeps
scripts/train-vae.sh from running at the reference's HEAD."""
import sys

import torch

from .config import Config
from .transformer import TransformerConfig
from .utils import to_device_i32
from ..engine import VAEConfig, VAEEngine
from ..MIDIUtil.defaults import *  # noqa: F401,F403
from ..MIDIUtil.defaults import SOS_ID


class LSTMConfig(Config):
    def __init__(self, n_layers: int, hidden_dim: int, dropout: float):
        super().__init__()
        self.n_layers = n_layers
        self.hidden_dim = hidden_dim
        self.dropout = dropout


class DecoderConfig(Config):
    def __init__(self, latent_dim: int, num_classes: int, output_dim: int,
                 transformer_config: TransformerConfig = None, lstm_config: LSTMConfig = None):
        super().__init__()
        assert (transformer_config is None) != (lstm_config is None), "give exactly one of transformer_config / lstm_config"
        self.transformer_config = transformer_config
        self.lstm_config = lstm_config
        self.latent_dim = latent_dim
        self.num_classes = num_classes
        self.output_dim = output_dim


class EncoderConfig(Config):
    def __init__(self, transformer_config: TransformerConfig, latent_dim: int, num_classes: int, input_dim: int):
        super().__init__()
        self.transformer_config = transformer_config
        self.latent_dim = latent_dim
        self.num_classes = num_classes
        self.input_dim = input_dim


class ModelConfig(Config):
    def __init__(self, encoder_config: EncoderConfig, decoder_config: DecoderConfig):
        super().__init__()
        self.encoder_config = encoder_config
        self.decoder_config = decoder_config


def to_engine_config(config: ModelConfig, featurisation: str = "events") -> VAEConfig:
    e, d = config.encoder_config, config.decoder_config
    et = e.transformer_config
    kw = dict(featurisation=featurisation, vocab=e.input_dim, num_classes=e.num_classes, enc_size=et.model_size, enc_layers=et.num_layers,
              enc_heads=et.num_heads, latent=e.latent_dim, enc_dropout=et.dropout)
    if d.lstm_config is not None:
        kw.update(dec_type="lstm", dec_size=d.lstm_config.hidden_dim, dec_layers=d.lstm_config.n_layers,
                  dec_dropout=d.lstm_config.dropout)
    else:
        dt = d.transformer_config
        kw.update(dec_type="transformer", dec_size=dt.model_size, dec_layers=dt.num_layers, dec_heads=dt.num_heads,
                  dec_dropout=dt.dropout)
    return VAEConfig(**kw)


class DecoderState:
    """Inference bookkeeping (model.py:107-128); the KV / value-sum caches live in the engine."""

    def __init__(self, batch_size: int, num_cache_layers: int, initial_state):
        self.reset(batch_size, num_cache_layers)
        self.initial_state = initial_state

    def advance_state(self, tokens):
        self.tokens = torch.cat([self.tokens, tokens.reshape(-1, 1).to(self.tokens.dtype)], dim=1)
        self.t += 1

    def reset(self, batch_size: int, num_cache_layers: int):
        self.tokens = torch.full((batch_size, 1), float(SOS_ID))
        self.t = 1
        self.caches = [{} for _ in range(num_cache_layers)]


class _EngineView:
    def __init__(self, engine: VAEEngine, prefix: str):
        self.engine = engine
        self._prefix = prefix

    def collect_params(self):
        return {k: self.engine.arena.view(k) for k in self.engine.arena.names() if k.startswith(self._prefix)}


class Encoder(_EngineView):
    """model.py:57-104: (tokens [B,T], seq_length [B], classes [B]) -> (means, stddevs), each [B, latent]."""

    def __init__(self, config: EncoderConfig, engine: VAEEngine):
        super().__init__(engine, "encoder.")
        self.config = config

    def __call__(self, tokens, seq_length, classes):
        dev = self.engine.device
        return self.engine.encode(to_device_i32(tokens, dev), to_device_i32(classes, dev))


class _DecoderBase(_EngineView):
    def __init__(self, config: DecoderConfig, engine: VAEEngine):
        super().__init__(engine, "decoder.")
        self.config = config

    def forward_train(self, F, tokens, seq_length, hidden_states, classes):
        """(tokens, seq_length, latent z [B, latent], classes) -> probs [B, T, vocab] (model.py:172-183 / :237-257)."""
        dev = self.engine.device
        out = self.engine.forward(to_device_i32(tokens, dev), to_device_i32(seq_length, dev), to_device_i32(classes, dev),
                                  None, train=False, want_probs=True,
                                  z_override=torch.as_tensor(hidden_states).to(dev, torch.float32).contiguous())
        return out["probs"]

    def __call__(self, tokens, seq_length, hidden_states, classes):
        return self.forward_train(None, tokens, seq_length, hidden_states, classes)


class LSTMDecoder(_DecoderBase):
    """model.py:131-203."""

    def get_initial_state(self, F, classes, hidden_state):
        """latent2hid(z) + class2hid[classes], repeated per layer and split into (h0, c0) (model.py:159-167)."""
        eng = self.engine
        z = torch.as_tensor(hidden_state).to(eng.device, torch.float32)
        t = eng.decoder_initial_state(to_device_i32(classes, eng.device), z)
        H = eng.cfg.dec_size
        n = eng.cfg.dec_layers
        return [t[:, :H].unsqueeze(0).repeat(n, 1, 1), t[:, H:].unsqueeze(0).repeat(n, 1, 1)]


class Decoder(_DecoderBase):
    """model.py:206-272 (Transformer decoder)."""

    def get_initial_state(self, F, classes, hidden_state):
        eng = self.engine
        z = torch.as_tensor(hidden_state).to(eng.device, torch.float32)
        return eng.decoder_initial_state(to_device_i32(classes, eng.device), z).unsqueeze(1)


class Model:
    """model.py:275-296.  ``context`` may be None / 'gpu' / a torch device; there is no CPU context."""

    def __init__(self, config: ModelConfig, context=None, precision="bf16p3f", seed=0, quiet=False, featurisation="events",
                 *args, **kwargs):
        if not quiet:
            print("Creating a model with the following configuration:")
            config.output_to_stream(sys.stdout)
        self.config = config
        self.engine_seed = seed                    # --seed: also the seed of initialize() (Trainer._initialize_model)
        device = context if isinstance(context, (torch.device, str)) and str(context).startswith("cuda") else \
            torch.device("cuda", torch.cuda.current_device())
        self.engine = VAEEngine(to_engine_config(config, featurisation), device, seed=seed, precision=precision)
        dcls = LSTMDecoder if config.decoder_config.lstm_config is not None else Decoder
        self.decoder = dcls(config.decoder_config, self.engine)
        self.encoder = Encoder(config.encoder_config, self.engine)

    def __call__(self, tokens, seq_lens, classes, eps=None):
        """-> (probs [B,T,V], means [B,Z], vars [B,Z]); z = means + N(0,1) * vars (model.py:292)."""
        dev = self.engine.device
        out = self.engine.forward(to_device_i32(tokens, dev), to_device_i32(seq_lens, dev), to_device_i32(classes, dev),
                                  None, eps=eps, train=False, want_probs=True)
        return out["probs"], out["means"], out["stds"]

    hybrid_forward = __call__

    def collect_params(self):
        return {k: self.engine.arena.view(k) for k in self.engine.arena.names()}

    def initialize(self, init=None, ctx=None, seed=None):
        """mx.init.Xavier() (trainer.py:103-105)."""
        self.engine.arena.init_xavier(self.engine_seed if seed is None else seed)

    def hybridize(self, *a, **k):
        pass

    def save_parameters(self, path):
        torch.save(self.engine.arena.state_dict(), path)

    def load_parameters(self, path, ctx=None):
        self.engine.arena.load_state(torch.load(path, map_location="cpu"))
