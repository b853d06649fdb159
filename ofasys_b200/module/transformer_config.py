"""Config dataclasses of the transformer stack (fields the hot path reads from
ofasys/module/transformer_config.py:22-170)."""
import re
from dataclasses import dataclass, field
from typing import Optional

from ..configure import BaseDataclass

DEFAULT_MAX_SOURCE_POSITIONS = 1024
DEFAULT_MAX_TARGET_POSITIONS = 1024
_NAME_PARSER = r"(decoder|encoder)_(.*)"


@dataclass
class EncDecBaseConfig(BaseDataclass):
    embed_dim: Optional[int] = 512
    ffn_embed_dim: int = 2048
    layers: int = 6
    attention_heads: int = 8
    normalize_before: bool = False
    learned_pos: bool = False
    layerdrop: float = 0.0


@dataclass
class DecoderConfig(EncDecBaseConfig):
    input_dim: Optional[int] = None
    output_dim: Optional[int] = None

    def __post_init__(self):
        if self.input_dim is None:
            self.input_dim = self.embed_dim
        if self.output_dim is None:
            self.output_dim = self.embed_dim


@dataclass
class TransformerConfig(BaseDataclass):
    activation_fn: str = "relu"
    dropout: float = 0.1
    attention_dropout: float = 0.0
    activation_dropout: float = 0.0
    encoder: EncDecBaseConfig = field(default_factory=EncDecBaseConfig)
    max_source_positions: int = DEFAULT_MAX_SOURCE_POSITIONS
    decoder: DecoderConfig = field(default_factory=DecoderConfig)
    max_target_positions: int = DEFAULT_MAX_TARGET_POSITIONS
    share_decoder_input_output_embed: bool = False
    share_all_embeddings: bool = False
    layernorm_embedding: bool = False
    no_scale_embedding: bool = False
    no_cross_attention: bool = False
    cross_self_attention: bool = False
    checkpoint_activations: bool = False
    offload_activations: bool = False

    # flat-namespace access (`cfg.encoder_embed_dim`), as the reference's __getattr__/__setattr__
    def __getattr__(self, name):
        m = re.match(_NAME_PARSER, name)
        if m and m[1] in self.__dict__:
            return getattr(self.__dict__[m[1]], m[2])
        raise AttributeError(f"invalid argument {name}.")

    def __setattr__(self, name, value):
        m = re.match(_NAME_PARSER, name)
        if m and m[1] in self.__dict__:
            setattr(self.__dict__[m[1]], m[2], value)
        else:
            super().__setattr__(name, value)
