from .ofa import GeneralistModel, GeneralistModelConfig, OFAEncoderDecoderExecutor
from .transformer import TransformerDecoder, TransformerEncoder

BaseModel = GeneralistModel.__mro__[1]
__all__ = ["GeneralistModel", "GeneralistModelConfig", "OFAEncoderDecoderExecutor", "TransformerEncoder", "TransformerDecoder"]
