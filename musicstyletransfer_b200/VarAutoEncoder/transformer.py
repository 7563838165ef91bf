"""TransformerConfig (transformer.py:8-21).  The layer arithmetic of the reference's transformer.py
(DualFeedForward, MultiHeadDotAttention, Transformer{Encoder,Decoder}Layer, positional_encodings) is
executed by the kernels behind musicstyletransfer_b200.engine.VAEEngine."""
from typing import Optional

from .config import Config
from ..engine import positional_encodings  # noqa: F401  (transformer.py:204-211, same table)


class TransformerConfig(Config):
    def __init__(self, model_size: int, dropout: float, num_layers: int, num_heads: int,
                 vocab_size: Optional[int] = None):
        super().__init__()
        self.model_size = model_size
        self.dropout = dropout
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.vocab_size = vocab_size
