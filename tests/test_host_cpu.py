"""CPU (no GPU): host-side logic of the product -- registry, state-dict contract, integer position
machinery (bit-exact vs the reference's own tables), C-ABI export table, loud failure without CUDA."""
import ctypes
import hashlib
import os
import re

import pytest
import torch

import ofasys_b200 as ob
from ofasys_b200 import _lib
from ofasys_b200.adaptor.audio import make_audio_bucket_1d
from ofasys_b200.adaptor.text import make_token_bucket_position
from oracle import cases
from oracle import oracle_model as om

from util import TTS_ONLY, build_product, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["text_A", "text_B", "patch_B", "audio_A", "resnet_A", "video_A", "cfg1_tiny"])
def test_state_dict_contract(name):
    """Parameter names / shapes equal the reference's (golden `spec` was dumped from its state_dict)."""
    g = load_golden(name)
    m = build_product(name)
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.is_floating_point() and not k.endswith(".version")}
    ref = {k: v for k, v in g["spec"].items() if not any(t in k.split(".") for t in TTS_ONLY)}
    assert set(ours) == set(ref), (sorted(set(ours) - set(ref))[:5], sorted(set(ref) - set(ours))[:5])
    for k in ref:
        assert ours[k] == tuple(ref[k]), k
    # reference state dicts (incl. its derived buffers / TTS-only tensors) load
    sd = cases.synth_state_dict(g["spec"], seed=0)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith("rp_bucket") or k.endswith("version") for k in missing), missing
    # tied embedding is one tensor (general.py:191-221)
    assert m.encoder.adaptor.embed_tokens.weight.data_ptr() == m.decoder.adaptor.embed_tokens.weight.data_ptr()


def test_registry():
    st = ob.ConfigStore()
    assert st.get("ofasys.model", "unify").target is ob.GeneralistModel
    for n in ("text", "image_patch_embed", "audio_fbank"):
        assert st.contain("ofasys.adaptor", n)


def test_token_bucket_bit_exact():
    g = load_golden("text_A")
    shape, dig, corner = g["ints"]["encoder.adaptor.text.token_rp_bucket"]
    t = make_token_bucket_position(256, 1024)
    assert t.dtype == torch.int64 and tuple(t.shape) == shape
    assert hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest() == dig
    m = build_product("text_A")
    assert torch.equal(m.encoder.adaptor.text.token_rp_bucket, t)
    assert torch.equal(m.encoder.adaptor.text.rel_idx(24).long(), t[:24, :24])


def test_audio_bucket_bit_exact():
    """1-D (i - j) table reproduces the reference's 4096 x 4096 buffer bit for bit."""
    g = load_golden("audio_A")
    shape, dig, corner = g["ints"]["encoder.adaptor.audio_fbank.audio_rp_bucket"]
    one = make_audio_bucket_1d(1024)
    i = torch.arange(4096)
    full = one[i[:, None] - i[None, :] + 4095]
    assert tuple(full.shape) == shape
    assert hashlib.sha256(full.contiguous().numpy().tobytes()).hexdigest() == dig
    m = build_product("audio_A")
    assert torch.equal(m.encoder.adaptor.audio_fbank.rel_idx(49).long(), full[:49, :49])


def test_audio_lengths_and_box_bins():
    from ofasys_b200.adaptor.audio import Conv2dSubsampling4
    from ofasys_b200.preprocessor import quantize_box

    s = Conv2dSubsampling4(80, 8)
    lens = torch.tensor([700, 998, 200, 150, 3, 4])
    assert torch.equal(s.get_out_seq_lens_tensor(lens), om.audio_out_lengths(lens))
    x = torch.tensor([0.0, 255.6, 511.0, 512.0, 100.25])
    assert torch.equal(quantize_box(x), om.quantize_box(x))


def test_global_rel_idx_block_diagonal():
    """concat(): slot-local bucket ids are offset into the concatenated table; -1 elsewhere."""
    m = build_product("audio_A")
    ga = m.encoder.adaptor
    from ofasys_b200.adaptor.base import AdaptorOutput

    a = ga.audio_fbank.rel_idx(5)
    t = ga.text.rel_idx(3)
    outs = [
        AdaptorOutput(torch.zeros(1, 5, 8), torch.zeros(1, 5, dtype=torch.bool), None, None, rel_idx=a, rel_tables=[torch.zeros(2047, 2)]),
        AdaptorOutput(torch.zeros(1, 3, 8), torch.zeros(1, 3, dtype=torch.bool), None, None, rel_idx=t, rel_tables=[torch.zeros(511, 2)]),
    ]
    ga.COMPACT_ABOVE = 1 << 30
    idx, used = ga._global_idx(outs)
    assert used is None and idx.shape == (8, 8) and idx.dtype == torch.int32
    assert torch.equal(idx[:5, :5], a) and torch.equal(idx[5:, 5:], t + 2047)
    assert (idx[:5, 5:] == -1).all() and (idx[5:, :5] == -1).all()
    # compaction: ids renumbered to the rows this shape can address; `used` maps them back (bit-exact)
    ga.COMPACT_ABOVE = 1024
    ga._idx_cache.clear()
    cidx, used = ga._global_idx(outs)
    assert used is not None and used.numel() == len(set(idx[idx >= 0].tolist())) and cidx.dtype == torch.int32
    back = torch.where(cidx >= 0, used[cidx.clamp_min(0).long()].to(torch.int32), cidx)
    assert torch.equal(back, idx)


def test_cabi_exports_every_declared_symbol():
    """libofab.so loads without a GPU and exports exactly what include/ofab.h declares."""
    hdr = open(os.path.join(ROOT, "include", "ofab.h")).read()
    declared = set(re.findall(r"\b(ofab_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert os.path.exists(_lib.LIB_PATH), "libofab.so not built (run __graft_entry__.build())"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(L, sym), f"{sym} declared in ofab.h but not exported"
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    lib = _lib.lib()
    assert lib.ofab_version() >= 100
    assert lib.ofab_ln_partial_rows() == 296


def test_no_cpu_fallback():
    """The product refuses CPU tensors instead of silently computing elsewhere."""
    from ofasys_b200 import ops

    x = torch.randn(4, 64)
    w = torch.ones(64, dtype=torch.bfloat16)
    with pytest.raises(_lib.OfabError):
        ops.layer_norm(x, w, w)


def test_product_never_imports_oracle():
    import subprocess
    import sys

    code = "import sys, ofasys_b200, ofasys_b200.ops; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dp, _, fs in os.walk(os.path.join(ROOT, "ofasys_b200")):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_fused_adam_refuses_cpu_parameters():
    """The optimizer step has no CPU fallback either."""
    import ofasys_b200 as ob
    from ofasys_b200._lib import OfabError

    p = torch.nn.Parameter(torch.zeros(8, dtype=torch.bfloat16))
    with pytest.raises(OfabError):
        ob.FusedAdam([p])


def test_new_ops_refuse_cpu_tensors():
    """Criteria, dropout and the audio front end have no CPU fallback either: they raise instead of computing."""
    from ofasys_b200 import ops
    from ofasys_b200._lib import OfabError
    from ofasys_b200.preprocessor.audio import Fbank, utterance_cmvn_

    x = torch.zeros(4, 3, 16, dtype=torch.float32)
    with pytest.raises(OfabError):
        ops.ctc_loss_sum(x, torch.zeros(4, 2, dtype=torch.long), torch.ones(4, dtype=torch.long))
    with pytest.raises(OfabError):
        Fbank()(torch.zeros(1, 16000))
    with pytest.raises(OfabError):
        utterance_cmvn_(torch.zeros(1, 10, 80))
    with pytest.raises(OfabError):
        ops.cross_entropy_sum(torch.zeros(4, 16, dtype=torch.bfloat16), torch.zeros(4, dtype=torch.long), label_smoothing=0.1)


def test_incremental_state_keys_are_per_module():
    """Two attention modules never share a cache entry in the caller's incremental_state dict."""
    from ofasys_b200.module import MultiheadAttention

    a = MultiheadAttention(128, 2, self_attention=True)
    b = MultiheadAttention(128, 2, self_attention=True)
    assert a._state_key != b._state_key and a._state_key == a._state_key
    st = {}
    assert a.reorder_incremental_state(st, torch.tensor([0])) is st  # nothing cached yet: a no-op


def test_ctypes_binding_matches_the_header():
    """The ctypes table (_lib._SIGS / Structures) and include/ofab.h describe the same ABI: argument counts of every
    prototype, and sizeof of every struct as a C compiler lays it out."""
    import shutil
    import subprocess
    import tempfile

    hdr = open(os.path.join(ROOT, "include", "ofab.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|const char\*)\s+(ofab_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", code, flags=re.S)
    assert len(protos) >= 50
    for name, args in protos:
        args = args.strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        assert n == len(_lib._SIGS[name][1]), (name, n, len(_lib._SIGS[name][1]))
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"ofab_dropout": _lib.Dropout, "ofab_attn_fwd_args": _lib.AttnFwdArgs, "ofab_attn_bwd_args": _lib.AttnBwdArgs,
               "ofab_embed_ln_args": _lib.EmbedLnArgs, "ofab_embed_ln_bwd_args": _lib.EmbedLnBwdArgs, "ofab_ctc_args": _lib.CtcArgs,
               "ofab_adam_tensor": _lib.AdamTensor, "ofab_adam_hyper": _lib.AdamHyper, "ofab_ce_rows_args": _lib.CeRowsArgs}
    assert set(structs) == set(re.findall(r"}\s*(ofab_[a-z_]+)\s*;", code))  # every typedef'd struct of the header is bound
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "s.c")
        with open(src, "w") as f:
            f.write('#include <stdio.h>\n#include "ofab.h"\nint main(void){\n')
            for s in structs:
                f.write(f'printf("{s} %zu\\n", sizeof({s}));\n')
            f.write("return 0;}\n")
        exe = os.path.join(td, "s")
        subprocess.run([cc, "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for s, cls in structs.items():
        assert int(sizes[s]) == ctypes.sizeof(cls), (s, sizes[s], ctypes.sizeof(cls))


def test_bucket_map_packing_for_the_attention_kernels():
    """ops._idx16: int32 [Tq, Tk] bucket map -> int16 rows padded to an even length with -1, plus the transposed copy
    the dK/dV kernel gathers from (host logic, device-agnostic)."""
    from ofasys_b200 import ops

    g = torch.Generator().manual_seed(0)
    for Tq, Tk in ((5, 7), (8, 8), (1, 13), (33, 64)):
        idx = torch.randint(-1, 2047, (Tq, Tk), generator=g, dtype=torch.int32)
        a, b = ops._idx16(idx)
        assert a.dtype == torch.int16 and b.dtype == torch.int16
        assert a.shape == (Tq, (Tk + 1) // 2 * 2) and b.shape == (Tk, (Tq + 1) // 2 * 2)
        assert torch.equal(a[:, :Tk].int(), idx) and torch.equal(b[:, :Tq].int(), idx.t())
        assert bool((a[:, Tk:] == -1).all()) and bool((b[:, Tq:] == -1).all())
        a2, b2 = ops._idx16(idx)
        assert a2 is a and b2 is b  # cached per map object
