"""TEST INFRASTRUCTURE ONLY -- import shim that loads the UNMODIFIED reference hot-path files.

Only `oracle/make_golden.py` and `tests/test_oracle_vs_reference.py` (skipped when
/root/reference is absent, i.e. on the GPU box) import this.  The product path never does.

The reference package cannot be imported as a whole here (missing oss2/omegaconf/hydra/...,
and Python >= 3.11 rejects its mutable dataclass defaults), but its model/module/adaptor files
load once the few unrelated imports are stubbed (recipe: SURVEY.md Appendix A).  Nothing of the
reference is copied: the files are executed from where they lie under /root/reference.
"""
import copy
import dataclasses
import importlib
import os
import sys
import types
from contextlib import nullcontext
from enum import Enum, unique

REF_ROOT = os.environ.get("OFASYS_REFERENCE", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "ofasys")


def available():
    return os.path.isdir(REF_PKG)


_installed = False


def _patch_dataclasses():
    # restore <=3.10 behaviour: a dataclass *instance* as a field default is allowed
    orig = dataclasses._get_field

    def _get_field(cls, a_name, a_type, default_kw_only):
        d = getattr(cls, a_name, dataclasses.MISSING)
        if dataclasses.is_dataclass(d) and not isinstance(d, type):
            setattr(cls, a_name, dataclasses.field(default_factory=lambda p=d: copy.deepcopy(p)))
        elif (
            isinstance(d, dataclasses.Field)
            and dataclasses.is_dataclass(d.default)
            and not isinstance(d.default, type)
        ):
            proto = d.default
            d.default = dataclasses.MISSING
            d.default_factory = lambda p=proto: copy.deepcopy(p)
        return orig(cls, a_name, a_type, default_kw_only)

    dataclasses._get_field = _get_field


def _mod(name, path=None, **attrs):
    m = types.ModuleType(name)
    if path is not None:
        m.__path__ = path
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install(adaptors=("text", "image_patch_embed", "image_resnet", "audio", "video_image_sequence")):
    """Install the stub modules; returns the namespace of reference classes."""
    global _installed
    if _installed:
        return _namespace()
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    _patch_dataclasses()

    # --- omegaconf stub -------------------------------------------------------------------
    _mod(
        "omegaconf",
        II=lambda s: "${%s}" % s,
        DictConfig=dict,
        OmegaConf=type("OmegaConf", (), {}),
        open_dict=nullcontext,
        MISSING="???",
    )

    # --- root package: real __path__, no heavy __init__ -----------------------------------
    @unique
    class ModalityType(Enum):  # mirrors ofasys/__init__.py:28-45 (enum values are the contract)
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

    _mod("ofasys", path=[REF_PKG], ModalityType=ModalityType)

    # --- ofasys.configure stub (registry semantics of config_store.py:22-70,203-227) -------
    @dataclasses.dataclass
    class BaseDataclass:
        _name: str = None

        @classmethod
        def from_namespace(cls, args):
            if isinstance(args, cls):
                return args
            cfg = cls()
            for k in cfg.__dataclass_fields__:
                if not k.startswith("_") and hasattr(args, k):
                    setattr(cfg, k, getattr(args, k))
            return cfg

    class _Node:
        def __init__(self, target, config=None):
            self.target, self.config, self.is_active = target, config, False

    class ConfigStore:
        _inst = None

        def __new__(cls):
            if cls._inst is None:
                cls._inst = super().__new__(cls)
                cls._inst.repo = {}
            return cls._inst

        def store(self, group, name, obj, dc=None):
            self.repo[f"{group}.{name}"] = _Node(obj, dc() if dc is not None else None)

        def get(self, group, name=None):
            return self.repo[f"{group}.{name}"]

        def contain(self, group, name):
            return f"{group}.{name}" in self.repo

        def make_dataclass(self, group_key, dc_name, module_name, prefix_names=()):
            prefix_names = list(prefix_names)

            def cmp(val):
                name = val[0].rsplit(".", 1)[1]
                if name in prefix_names:
                    return (prefix_names.index(name), "")
                return (len(prefix_names), val[0])

            flds = []
            for path, node in sorted(self.repo.items(), key=cmp):
                group, name = path.rsplit(".", 1)
                if group == group_key and node.config is not None:
                    flds.append((name, type(node.config), dataclasses.field(default_factory=node.config.__class__)))
            c = dataclasses.make_dataclass(dc_name, flds, bases=(BaseDataclass,))
            c.__module__ = module_name
            return c

    def register_config(group, name, dataclass=None):
        def _r(cls):
            ConfigStore().store(group, name, cls, dataclass)
            return cls

        return _r

    wanted = list(adaptors)

    def auto_import(init_file):
        # import only the adaptor modules on the hot path (general.py:22 auto-imports all)
        if os.path.basename(os.path.dirname(init_file)) == "adaptor":
            for a in wanted:
                importlib.import_module(f"ofasys.adaptor.{a}")

    cfgmod = _mod(
        "ofasys.configure",
        path=[],
        BaseDataclass=BaseDataclass,
        ChoiceEnum=lambda choices: str,
        register_config=register_config,
        ConfigStore=ConfigStore,
        auto_import=auto_import,
    )
    _mod("ofasys.configure.utils", convert_namespace_to_omegaconf=lambda a: a, gen_parser_from_dataclass=lambda *a, **k: None)
    _mod("ofasys.configure.config_store", register_config=register_config, ConfigStore=ConfigStore)
    _mod("ofasys.configure.configs", BaseDataclass=BaseDataclass)
    cfgmod.utils = sys.modules["ofasys.configure.utils"]

    # --- distributed / utils stubs ----------------------------------------------------------
    du = _mod("ofasys.distributed.utils")
    _mod("ofasys.distributed", path=[], fsdp_wrap=lambda m, **k: m, utils=du)
    _mod("ofasys.module.fused_kernels", _is_fused_kernel_available=False)
    _mod("ofasys.utils", path=[])
    _mod("ofasys.utils.file_utils", cached_path=lambda p, *a, **k: p)
    _mod("ofasys.utils.logging_utils", master_logging=lambda *a, **k: (lambda f: f))

    # --- preprocessor: real dir for instruction.py, minimal Dictionary ----------------------
    class Dictionary:
        """6-method stand-in for preprocessor/dictionary.py:20-48 (<s>=0 <pad>=1 </s>=2 <unk>=3)."""

        def __init__(self, n):
            self.n = n
            self.indices = {}

        def __len__(self):
            return self.n

        def bos(self):
            return 0

        def pad(self):
            return 1

        def eos(self):
            return 2

        def unk(self):
            return 3

    pp = _mod("ofasys.preprocessor", path=[os.path.join(REF_PKG, "preprocessor")], Dictionary=Dictionary)
    _mod("ofasys.preprocessor.dictionary", Dictionary=Dictionary)
    instr = importlib.import_module("ofasys.preprocessor.instruction")
    pp.Slot = instr.Slot
    pp.Instruction = instr.Instruction

    _installed = True
    return _namespace()


def _namespace():
    ofa = importlib.import_module("ofasys.model.ofa")
    ns = types.SimpleNamespace()
    ns.GeneralistModel = ofa.GeneralistModel
    ns.GeneralistModelConfig = ofa.GeneralistModelConfig
    ns.Slot = sys.modules["ofasys.preprocessor"].Slot
    ns.Dictionary = sys.modules["ofasys.preprocessor"].Dictionary
    ns.ModalityType = sys.modules["ofasys"].ModalityType
    ns.text = importlib.import_module("ofasys.adaptor.text")
    return ns


def build_reference_model(
    arch="tiny",
    enc_layers=None,
    dec_layers=None,
    vocab=50265,
    adaptors=("text",),
    mode="A",
    seed=0,
    resnet_drop_path_rate=0.0,
    dims=None,
    resnet_type=None,
):
    """Build the reference GeneralistModel with default_model.yaml settings, dropout 0.

    mode "A": use_self_attn_bias=True, positions disentangled (released OFA+ checkpoints).
    mode "B": use_self_attn_bias=False, entangle_position_embedding=True (needed by
              image_patch_embed).
    """
    import torch

    ns = install()
    cfg = ns.GeneralistModelConfig()
    cfg.arch = arch
    cfg.activation_fn = "gelu"
    cfg.dropout = 0.0
    cfg.attention_dropout = 0.0
    cfg.share_all_embeddings = True
    cfg.share_decoder_input_output_embed = True
    cfg.no_scale_embedding = True
    cfg.layernorm_embedding = True
    cfg.encoder.normalize_before = True
    cfg.decoder.normalize_before = True
    cfg.encoder.learned_pos = True
    cfg.decoder.learned_pos = True
    if mode == "B":
        cfg.use_self_attn_bias = False
        cfg.entangle_position_embedding = True
    m = ns.GeneralistModel(cfg)
    if dims is not None:  # (embed_dim, heads, ffn_dim): override the arch preset (ofa.py:352-354)
        d, h, f = dims
        m.cfg.encoder.embed_dim = m.cfg.decoder.embed_dim = d
        m.cfg.encoder.ffn_embed_dim = m.cfg.decoder.ffn_embed_dim = f
        m.cfg.decoder.input_dim = m.cfg.decoder.output_dim = d
        m.cfg.encoder.attention_heads = m.cfg.decoder.attention_heads = h
    if enc_layers is not None:
        m.cfg.encoder.layers = enc_layers
    if dec_layers is not None:
        m.cfg.decoder.layers = dec_layers
    for f in dataclasses.fields(m.cfg.adaptor):
        if f.name.startswith("_"):
            continue
        a = getattr(m.cfg.adaptor, f.name)
        name = f.name
        a.is_active = name in adaptors or (name == "audio_fbank" and "audio" in adaptors)
        if mode == "B":
            a.entangle_position_embedding = True
        if hasattr(a, "drop_path_rate"):
            a.drop_path_rate = resnet_drop_path_rate
    if "image_patch_embed" in adaptors:
        m.cfg.adaptor.image_patch_embed.embed_dim = m.cfg.encoder.embed_dim
    if resnet_type is not None:  # the arch presets pick it (ofa.py:557-650: base resnet101, large resnet152)
        m.cfg.adaptor.image_resnet.resnet_type = resnet_type
    torch.manual_seed(seed)
    m.initialize(ns.Dictionary(vocab))
    return m, ns
