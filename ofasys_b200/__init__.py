"""ofasys_b200 -- B200-native (sm_100a) implementation of the OFASys unified multimodal
encoder-decoder hot path behind the reference's GeneralistModel / Slot / adaptor-plugin API.
See DESIGN.md (scope, kernels, rooflines) and INTEGRATION.md (how it drops under ofasys.Trainer)."""
from .configure import BaseDataclass, ConfigStore, register_config
from .preprocessor import Dictionary, ModalityType, Slot
from .model import GeneralistModel, GeneralistModelConfig
from .optim import FusedAdam
from .criterion import LabelSmoothedCrossEntropyCriterion, SpeechToTextLossCriterion

__all__ = ["ModalityType", "Slot", "Dictionary", "GeneralistModel", "GeneralistModelConfig", "BaseDataclass", "ConfigStore", "register_config", "FusedAdam",
           "LabelSmoothedCrossEntropyCriterion", "SpeechToTextLossCriterion"]
