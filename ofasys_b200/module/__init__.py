from .initialize import init_bert_params
from .layer import Embedding, Linear
from .layer_norm import LayerNorm
from .resnet import resnet50_backbone, resnet101_backbone, resnet152_backbone
from .multihead_attention import MultiheadAttention
from .transformer_config import DecoderConfig, EncDecBaseConfig, TransformerConfig
from .transformer_layer import TransformerDecoderLayer, TransformerEncoderLayer

__all__ = [
    "Embedding", "Linear", "LayerNorm", "MultiheadAttention", "TransformerConfig", "EncDecBaseConfig", "DecoderConfig",
    "TransformerEncoderLayer", "TransformerDecoderLayer", "init_bert_params",
]
