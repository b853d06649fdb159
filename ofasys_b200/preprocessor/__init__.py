"""Data contract at the hot path's entry: Slot (ofasys/preprocessor/instruction.py:29-106),
ModalityType (ofasys/__init__.py:28-45), the Dictionary specials (preprocessor/dictionary.py:41-48)
and the integer box-bin quantisation (preprocessor/default/box.py:101-110)."""
from dataclasses import dataclass
from enum import Enum, unique
from typing import Any, Dict, List, Optional


@unique
class ModalityType(Enum):
    TEXT = 1
    IMAGE = 2
    BOX = 3
    AUDIO = 4
    MOTION = 5
    PHONE = 6
    VIDEO = 7
    STRUCT = 8
    CATEGORY = 9

    @classmethod
    def parse(cls, mark):
        for mod in ModalityType:
            if mark == mod.name:
                return cls(mod.value)
        return None


@dataclass
class Slot:
    modality: ModalityType
    is_src: bool
    value: Optional[Any]
    global_position: Optional[int] = None
    column_name: Optional[str] = None
    attributes: Optional[List[str]] = None
    preprocess: Optional[str] = None
    is_plaintext: bool = False
    split: str = "train"
    decoder_plain_with_loss: bool = False

    def __post_init__(self):
        if self.column_name is None:
            self.column_name = str(self.global_position)
        if self.attributes is not None and isinstance(self.attributes, str):
            self.attributes = self.attributes.split(",")

    def has_attr(self, attr_key: str) -> bool:
        return any(a == attr_key or a.startswith(attr_key + "=") for a in (self.attributes or []))

    def get_attr(self, attr_key: str, class_factory: type = None):
        for a in self.attributes or []:
            if a.startswith(attr_key + "="):
                val = a[len(attr_key) + 1:]
                return class_factory(val) if class_factory is not None else val
        return None

    @staticmethod
    def get_target_slot_from_slots(slots: List):
        return [s for s in slots if not s.is_src][-1]

    @staticmethod
    def get_target_slot_from_sample(sample: Dict):
        return [s for s in sample["net_input"]["slots"] if not s.is_src][-1]


class Dictionary:
    """Symbol table with the fairseq specials <s>=0 <pad>=1 </s>=2 <unk>=3."""

    def __init__(self, symbols=None, n_dummy: int = 0, num_bins: int = 0):
        self.symbols = ["<s>", "<pad>", "</s>", "<unk>"]
        self.indices = {s: i for i, s in enumerate(self.symbols)}
        for s in symbols or []:
            self.add_symbol(s)
        for i in range(n_dummy):
            self.add_symbol(f"madeupword{i:06d}")
        for i in range(num_bins):
            self.add_symbol(f"<bin>_{i}")

    def add_symbol(self, s):
        if s not in self.indices:
            self.indices[s] = len(self.symbols)
            self.symbols.append(s)
        return self.indices[s]

    def index(self, s):
        return self.indices.get(s, self.unk())

    def __len__(self):
        return len(self.symbols)

    def bos(self):
        return 0

    def pad(self):
        return 1

    def eos(self):
        return 2

    def unk(self):
        return 3


def quantize_box(coords, max_image_size: int = 512, num_bins: int = 1000):
    """coordinate -> bin index k of `<bin>_k` = round(x / max_image_size * (num_bins - 1)); integer-exact
    restatement of preprocessor/default/box.py:101-110 (torch.round = round-half-to-even, as there)."""
    return (coords / max_image_size * (num_bins - 1)).round().long()
