"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch, fp32) of the OFASys hot path.

This is the parity oracle: a function-by-function restatement of the reference's unified
encoder-decoder forward (ofasys/model + ofasys/module + ofasys/adaptor), written against the
reference's *parameter names* so the same state_dict drives the oracle, the reference (through
oracle/ref_shim.py, container only) and the CUDA product.  Gradients come from torch autograd on
CPU.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import it; the product (ofasys_b200/) never does and has no CPU fallback.

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by
the reference itself), so this restatement is pinned against outputs of the *reference code run
here* -- oracle/make_golden.py imports the unmodified reference files, dumps logits / loss / grad
statistics for seeded inputs into tests/golden/, and tests/test_oracle_golden.py replays them
through this file (rel-L2 <= 1e-5 fp32).  Each function cites the reference lines it follows
(paths relative to /root/reference/ofasys).
"""
import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import torch
import torch.nn.functional as F

TEXT, IMAGE, BOX, AUDIO, MOTION, PHONE, VIDEO, STRUCT, CATEGORY = range(1, 10)  # __init__.py:28-45
PAD = 1  # preprocessor/dictionary.py:41-48  (<s>=0 <pad>=1 </s>=2 <unk>=3)


@dataclass
class OracleConfig:
    """Subset of GeneralistModelConfig (model/ofa.py:41-122) that the hot path reads."""

    embed_dim: int = 256
    heads: int = 4
    ffn_dim: int = 1024
    enc_layers: int = 4
    dec_layers: int = 4
    vocab: int = 50265
    mode: str = "A"  # "A": use_self_attn_bias, disentangled pos. "B": no bias, entangled pos.
    attn_scale_factor: float = 2.0  # ofa.py:56
    token_bucket_size: int = 256  # adaptor/text.py:34
    max_position: int = 1024  # module/transformer_config.py:14
    image_bucket_size: int = 42  # adaptor/image_resnet.py:62
    patch: int = 14  # adaptor/image_patch_embed.py:23-27
    image_size: int = 224
    audio_feat_dim: int = 80  # adaptor/audio.py:66
    resnet_type: str = "resnet101"
    ln_eps: float = 1e-5  # module/layer_norm.py:27

    @property
    def head_dim(self):
        return self.embed_dim // self.heads


@dataclass
class OSlot:
    """Mirror of preprocessor/instruction.py:29-51 (only the fields the model reads)."""

    modality: int
    is_src: bool
    value: Any
    adaptor: Optional[str] = None  # the `adaptor=<name>` slot attribute (general.py:103-118)


# ------------------------------------------------------------------ integer position machinery
def make_token_bucket_position(bucket_size: int, max_position: int) -> torch.Tensor:
    """adaptor/text.py:20-30 (identical: audio.py:50-60, video_image_sequence.py:50-60). int64."""
    context_pos = torch.arange(max_position, dtype=torch.long)[:, None]
    memory_pos = torch.arange(max_position, dtype=torch.long)[None, :]
    rel = context_pos - memory_pos
    sign = torch.sign(rel)
    mid = bucket_size // 2
    abs_pos = torch.where((rel < mid) & (rel > -mid), mid - 1, torch.abs(rel))
    log_pos = torch.ceil(torch.log(abs_pos / mid) / math.log((max_position - 1) / mid) * (mid - 1)) + mid
    log_pos = log_pos.int()
    bucket = torch.where(abs_pos.le(mid), rel, log_pos * sign).long()
    return bucket + bucket_size - 1


def make_image_bucket_position(bucket_size: int, num_relative_distance: int) -> torch.Tensor:
    """adaptor/image_resnet.py:25-40. int64 [(bs*bs+1), (bs*bs+1)]."""
    ch = torch.arange(bucket_size)
    cw = torch.arange(bucket_size)
    coords = torch.stack(torch.meshgrid([ch, cw], indexing="ij"))
    flat = torch.flatten(coords, 1)
    rel = flat[:, :, None] - flat[:, None, :]
    rel = rel.permute(1, 2, 0).contiguous()
    rel[:, :, 0] += bucket_size - 1
    rel[:, :, 1] += bucket_size - 1
    rel[:, :, 0] *= 2 * bucket_size - 1
    idx = torch.zeros(size=(bucket_size * bucket_size + 1,) * 2, dtype=rel.dtype)
    idx[1:, 1:] = rel.sum(-1)
    idx[0, 0:] = num_relative_distance - 3
    idx[0:, 0] = num_relative_distance - 2
    idx[0, 0] = num_relative_distance - 1
    return idx


def quantize_box(coords: torch.Tensor, max_image_size: int = 512, num_bins: int = 1000) -> torch.Tensor:
    """preprocessor/default/box.py:101-110: coordinate -> <bin_k> index k = round(x/max*(bins-1))."""
    return (coords / max_image_size * (num_bins - 1)).round().long()


def audio_out_lengths(in_lens: torch.Tensor) -> torch.Tensor:
    """module/subsample.py:37-41: floor((L-1)/2+1) twice (reference quirk 7: one frame too many)."""
    out = in_lens.clone()
    for _ in range(2):
        out = ((out.float() - 1) / 2 + 1).floor().long()
    return out


# ------------------------------------------------------------------ dropout sites
# The reference draws its masks from torch's RNG (module/dropout.py:14-25 -> F.dropout; module/droppath.py:13-63),
# the CUDA path from a counter-based hash, so the two can only be compared with the SAME masks: a parity test
# installs DROP_HOOK(kind, x) -> multiplier tensor broadcastable to x (0 or 1/keep, drop-path folded in) and the
# oracle multiplies it in at exactly the places the reference calls its Dropout / DropPath modules.  kind:
#   "embed" (adaptor/base.py:181, x is B x T x C), "attn_probs" (multihead_attention.py:335, B*H x T x S),
#   "branch" (transformer_layer.py:181,203 / :433,466,489 followed by drop_path :87,:333, x is T x B x C),
#   "act" (transformer_layer.py:195,481, T x B x 4C).  None = eval mode / p = 0 (every golden fixture).
DROP_HOOK = None


def _drop(kind, x):
    return x if DROP_HOOK is None else x * DROP_HOOK(kind, x)


# bf16 activation STORAGE model ("parity mode" of the oracle).  The CUDA path computes every contraction with fp32
# accumulation and keeps the residual stream, LayerNorm statistics, softmax and reductions in fp32, but STORES GEMM operands
# and outputs in bf16 (the reference's own common.bf16 mode stores everything in bf16, trainer.py:215-219).  With
# STORE_BF16 the oracle -- arithmetic still fp32, same algorithm, same order of reference lines -- rounds exactly at those
# storage points: Linear / conv outputs, the LayerNorm outputs that feed a GEMM, q / k / v, the attention probabilities
# before P V, the attention output, the per-head-scaled out_proj weight, the fp16 position-bias tile, the logits.  What is
# left between the CUDA path and this oracle is accumulation order (and one-ulp bf16 flips it causes): rounding and
# defects become separable -- tests/test_model_gpu.py gates the logits at 1e-3 rel-L2 against it (BASELINE north_star).
# It also serves the ResNet: a ReLU network's gradient is discontinuous in forward perturbations (a pre-activation that
# crosses 0 flips a whole gradient element; ~0.3 % of masks per ReLU, 20-48 % per parameter over 49 ReLUs), so the
# ResNet gradients are compared under the same storage model.
STORE_BF16 = False


class _RoundBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def _st(x):
    return _RoundBf16.apply(x) if STORE_BF16 else x


class _RoundFp16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.half().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def _st_bias(b):
    """the additive position tile is stored in fp16 in the log2 domain (csrc/attn_tc.cu attn_bias_build_kernel)"""
    if not STORE_BF16 or b is None:
        return b
    log2e = 1.4426950408889634
    return _RoundFp16.apply(b * log2e) / log2e



# ------------------------------------------------------------------ small modules
def layer_norm(x, sd, prefix, eps=1e-5, st=False):
    """module/layer_norm.py:27-32 -> torch.nn.LayerNorm.  st: the output feeds a GEMM (stored bf16 under STORE_BF16)."""
    y = F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)
    return _st(y) if st else y


def linear(x, sd, prefix, st=True):
    """st: the CUDA path stores this GEMM's output in bf16 (fp32 when it lands on the residual stream)."""
    y = F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))
    return _st(y) if st else y


def gelu(x):
    """module/gelu.py:18-19: exact erf GELU in fp32."""
    return F.gelu(x.float()).type_as(x)


# ------------------------------------------------------------------ attention
def mha(sd, prefix, cfg: OracleConfig, query, key, kpm, attn_mask, attn_bias, fast_path):
    """module/multihead_attention.py:113-353.  query/key are T x B x C.

    fast_path=True restates the F.multi_head_attention_forward short-circuit (:155-186) taken by
    encoder self-attention when attn_bias is None: scale head_dim**-0.5 and *no* c_attn
    (SURVEY.md 3.6 quirk 1).  Otherwise the manual path (:188-353): scale (head_dim*2)**-0.5,
    additive bias, causal mask added after nan_to_num, -inf on padded keys, fp32 softmax, c_attn.
    """
    T, B, C = query.shape
    S = key.shape[0]
    H, dh = cfg.heads, cfg.head_dim
    q = linear(query, sd, prefix + ".q_proj")
    k = linear(key, sd, prefix + ".k_proj")
    v = linear(key, sd, prefix + ".v_proj")
    scaling = float(dh) ** -0.5 if fast_path else float(dh * cfg.attn_scale_factor) ** -0.5  # :55
    q = q * scaling
    q = q.contiguous().view(T, B * H, dh).transpose(0, 1)
    k = k.contiguous().view(S, B * H, dh).transpose(0, 1)
    v = v.contiguous().view(S, B * H, dh).transpose(0, 1)
    w = torch.bmm(q, k.transpose(1, 2))  # :308
    if attn_bias is not None:
        w = w + _st_bias(attn_bias)  # :311-312
    if attn_mask is not None:
        w = torch.nan_to_num(w)  # :314-317
        w = w + attn_mask.to(w.dtype).unsqueeze(0)  # buffered_future_mask is cast to the activations (transformer.py:537)
    if kpm is not None:
        w = w.view(B, H, T, S).masked_fill(kpm.unsqueeze(1).unsqueeze(2).to(torch.bool), float("-inf"))
        w = w.view(B * H, T, S)  # :319-326
    if STORE_BF16:
        # storage model of the fused kernel: P V consumes UNNORMALISED probabilities 2^(x - m), m = the row maximum rounded up
        # to an integer in the log2 domain (any integer shift gives the same bf16 rounding), stored in bf16; the row sum is
        # that of the unrounded values and divides the product afterwards
        x2 = w.float() * 1.4426950408889634
        m = torch.ceil(x2.amax(dim=-1, keepdim=True))
        m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
        pu = torch.exp2(x2 - m)
        l = pu.sum(dim=-1, keepdim=True)
        a = torch.bmm(_st(_drop("attn_probs", pu)), v) / l.clamp_min(1e-38)
    else:
        p = F.softmax(w.float(), dim=-1).type_as(w)  # :333-334 (module/utils.py:451)
        p = _drop("attn_probs", p)  # :335
        a = torch.bmm(p, v)  # :338
    a = _st(a.transpose(0, 1).contiguous().view(T, B, C))
    if not fast_path and (prefix + ".c_attn") in sd:
        if STORE_BF16:  # the CUDA path folds the per-head scale into out_proj's weight columns (exact in real arithmetic)
            w_eff = _st(sd[prefix + ".out_proj.weight"] * sd[prefix + ".c_attn"].repeat_interleave(dh).unsqueeze(0))
            return _st(F.linear(a, w_eff, sd.get(prefix + ".out_proj.bias")))
        a = a.view(T, B, H, dh)
        a = torch.einsum("tbhd,h->tbhd", a, sd[prefix + ".c_attn"])  # :342-345
        a = a.reshape(T, B, C)
    return linear(a, sd, prefix + ".out_proj")


def ffn(sd, prefix, x):
    """transformer_layer.py:188-207: fc1 -> gelu(fp32) -> ffn_layernorm (scale_fc) -> fc2."""
    x = gelu(linear(x, sd, prefix + ".fc1"))
    x = _drop("act", x)  # :195 / :481
    if (prefix + ".ffn_layernorm.weight") in sd:
        x = layer_norm(x, sd, prefix + ".ffn_layernorm", st=True)
    return linear(x, sd, prefix + ".fc2")


def encoder_layer(sd, prefix, cfg, x, kpm, self_attn_bias):
    """module/transformer_layer.py:132-209 (pre-LN, normformer extras on; dropout 0)."""
    residual = x
    x = layer_norm(x, sd, prefix + ".self_attn_layer_norm", st=True)
    x = mha(sd, prefix + ".self_attn", cfg, x, x, kpm, None, self_attn_bias, fast_path=self_attn_bias is None)
    if (prefix + ".attn_ln.weight") in sd:
        x = layer_norm(x, sd, prefix + ".attn_ln")
    x = residual + _drop("branch", x)  # :181, :87
    residual = x
    x = layer_norm(x, sd, prefix + ".final_layer_norm", st=True)
    x = ffn(sd, prefix, x)
    return residual + _drop("branch", x)  # :203, :87


def decoder_layer(sd, prefix, cfg, x, enc, enc_kpm, self_mask, self_kpm, self_bias, cross_bias):
    """module/transformer_layer.py:351-495.  Decoder self-attention always takes the manual path
    (transformer.py:476-477 passes False, not None, when biases are off)."""
    residual = x
    x = layer_norm(x, sd, prefix + ".self_attn_layer_norm", st=True)
    x = mha(sd, prefix + ".self_attn", cfg, x, x, self_kpm, self_mask, self_bias, fast_path=False)
    if (prefix + ".self_attn_ln.weight") in sd:
        x = layer_norm(x, sd, prefix + ".self_attn_ln")
    x = residual + _drop("branch", x)  # :433, :333
    residual = x
    x = layer_norm(x, sd, prefix + ".encoder_attn_layer_norm", st=True)
    x = mha(sd, prefix + ".encoder_attn", cfg, x, enc, enc_kpm, None, cross_bias, fast_path=False)
    if (prefix + ".cross_attn_ln.weight") in sd:
        x = layer_norm(x, sd, prefix + ".cross_attn_ln")
    x = residual + _drop("branch", x)  # :466, :333
    residual = x
    x = layer_norm(x, sd, prefix + ".final_layer_norm", st=True)
    x = ffn(sd, prefix, x)
    return residual + _drop("branch", x)  # :489, :333


# ------------------------------------------------------------------ adaptors
@dataclass
class AOut:
    """adaptor/base.py:19-54."""

    embed: torch.Tensor  # B x T x d
    masks: torch.Tensor  # B x T bool
    pos_embed: torch.Tensor  # B x T x d
    self_attn_bias: List[Optional[torch.Tensor]] = field(default_factory=list)  # per layer B x H x T x T


def _hook(sd, ap, cfg: OracleConfig, slot: OSlot, out: AOut, num_layers: int, rel_fn):
    """adaptor/base.py:152-191 (forward hook of every adaptor; embed_scale 1.0, dropout 0)."""
    embed = 1.0 * out.embed
    if cfg.mode == "B" and out.pos_embed is not None:
        embed = embed + out.pos_embed  # :170-171 (entangle_position_embedding)
    if slot.is_src:
        embed = embed + sd[ap + ".type_embedding.weight"].squeeze()  # :172-173
    embed = layer_norm(embed, sd, ap + ".layernorm_embedding")
    if out.pos_embed is not None:
        out.pos_embed = layer_norm(out.pos_embed, sd, ap + ".layernorm_position", st=True)
    out.embed = _drop("embed", embed)  # :181
    if not out.self_attn_bias and cfg.mode == "A":
        B, T = embed.shape[:2]
        out.self_attn_bias = []
        for idx in range(num_layers):
            values = rel_fn(idx, T)  # T x T x H
            out.self_attn_bias.append(values.unsqueeze(0).expand(B, -1, -1, -1).permute(0, 3, 1, 2))  # :242-258
    return out


def text_adaptor(sd, gp, cfg, slot, num_layers):
    """adaptor/text.py:106-127 (+ get_rel_pos_bias :101-104)."""
    ap = gp + ".text"
    tok = slot.value
    masks = tok.eq(PAD)
    B, T = tok.shape
    pos = F.embedding(torch.arange(T).unsqueeze(0).expand(B, T), sd[ap + ".embed_positions.weight"])
    emb = F.embedding(tok, sd[gp + ".embed_tokens.weight"], padding_idx=PAD)
    bucket = make_token_bucket_position(cfg.token_bucket_size, cfg.max_position)

    def rel(idx, T):
        return F.embedding(bucket[:T, :T], sd[f"{ap}.token_rel_pos_table_list.{idx}.weight"])

    return _hook(sd, ap, cfg, slot, AOut(emb, masks, pos, []), num_layers, rel)


def patch_embed_adaptor(sd, gp, cfg, slot, num_layers):
    """adaptor/image_patch_embed.py:62-80 (Conv2d k=s=patch, CLS, learned positions)."""
    ap = gp + ".image_patch_embed"
    img = slot.value
    B = img.shape[0]
    x = _st(F.conv2d(_st(img), sd[ap + ".proj.weight"], sd[ap + ".proj.bias"], stride=cfg.patch))
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((sd[ap + ".cls_token"].expand(B, -1, -1), x), dim=1)
    T = x.shape[1]
    masks = torch.zeros((B, T), dtype=torch.bool)
    pos = F.embedding(torch.arange(T).unsqueeze(0).expand(B, -1), sd[ap + ".embed_image_positions.weight"])
    # self_attn_bias=None: only legal in mode B (quirk 2: abstract get_rel_pos_bias would raise)
    assert cfg.mode == "B", "image_patch_embed needs use_self_attn_bias=False (SURVEY 3.6 quirk 2)"
    return _hook(sd, ap, cfg, slot, AOut(x, masks, pos, []), num_layers, None)


def audio_adaptor(sd, gp, cfg, slot, num_layers):
    """adaptor/audio.py:295-325 (src branch) + module/subsample.py:43-63."""
    ap = gp + ".audio_fbank"
    fbank, lens = slot.value["fbank"], slot.value["fbank_lengths"]
    x = fbank.unsqueeze(1)
    x = _st(F.relu(F.conv2d(x, sd[ap + ".subsample.conv.0.weight"], sd[ap + ".subsample.conv.0.bias"], stride=2)))
    x = _st(F.relu(F.conv2d(x, sd[ap + ".subsample.conv.2.weight"], sd[ap + ".subsample.conv.2.bias"], stride=2)))
    b, c, t, f = x.shape
    x = linear(x.transpose(1, 2).contiguous().view(b, t, c * f), sd, ap + ".subsample.out.0")
    out_lens = audio_out_lengths(lens)
    masks = torch.zeros((b, t), dtype=torch.bool)
    for i, l in enumerate(out_lens.tolist()):  # audio.py:307-310
        diff = l - t
        if diff < 0:
            masks[i, diff:] = True
    pos = F.embedding(torch.arange(t).unsqueeze(0).expand(b, t), sd[ap + ".embed_audio_positions.weight"])
    bucket = make_token_bucket_position(cfg.max_position, 4096)  # audio.py:50-60,229,235

    def rel(idx, T):
        return F.embedding(bucket[:T, :T], sd[f"{ap}.audio_rel_pos_table_list.{idx}.weight"])

    return _hook(sd, ap, cfg, slot, AOut(x, masks, pos, []), num_layers, rel)


# --- ResNet backbone (module/resnet.py:139-261; torchvision-style bottlenecks, layers 1-3 only)
_RESNET_LAYERS = {"resnet50": [3, 4, 6], "resnet101": [3, 4, 23], "resnet152": [3, 8, 36]}


def _bn(x, sd, p, training):
    """nn.BatchNorm2d, eps 1e-5, momentum .1; train mode uses batch statistics (quirk 13)."""
    return F.batch_norm(
        x, sd[p + ".running_mean"].clone(), sd[p + ".running_var"].clone(), sd[p + ".weight"], sd[p + ".bias"],
        training=training, momentum=0.1, eps=1e-5,
    )


def _bottleneck(x, sd, p, stride, has_down, training):
    """module/resnet.py:116-136 (conv1x1-bn-relu, conv3x3(stride)-bn-relu, conv1x1-bn, +id, relu)."""
    idt = x
    out = _st(F.relu(_bn(_st(F.conv2d(x, sd[p + ".conv1.weight"])), sd, p + ".bn1", training)))
    out = _st(F.relu(_bn(_st(F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1)), sd, p + ".bn2", training)))
    out = _bn(_st(F.conv2d(out, sd[p + ".conv3.weight"])), sd, p + ".bn3", training)
    if has_down:
        idt = _st(_bn(_st(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride)), sd, p + ".downsample.1", training))
    return _st(F.relu(out + idt))


def resnet_backbone(x, sd, p, kind, training=True):
    """module/resnet.py:235-246: conv1(7x7,s2) bn relu maxpool(3,s2,p1) layer1..layer3 -> stride 16, 1024 ch."""
    x = _st(F.relu(_bn(_st(F.conv2d(_st(x), sd[p + ".conv1.weight"], stride=2, padding=3)), sd, p + ".bn1", training)))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, nblocks in enumerate(_RESNET_LAYERS[kind]):
        for bi in range(nblocks):
            stride = 2 if (li > 0 and bi == 0) else 1
            x = _bottleneck(x, sd, f"{p}.layer{li + 1}.{bi}", stride, bi == 0, training)
    return x


def image_resnet_adaptor(sd, gp, cfg, slot, num_layers, training=True):
    """adaptor/image_resnet.py:116-202."""
    ap = gp + ".image_resnet"
    img = slot.value
    B = img.shape[0]
    feat = resnet_backbone(img, sd, ap + ".embed_images", cfg.resnet_type, training)
    h, w = feat.shape[-2:]
    P = h * w
    masks = torch.zeros((B, P), dtype=torch.bool)
    pid = (torch.arange(w).unsqueeze(0).expand(h, w) + torch.arange(h).unsqueeze(1) * cfg.image_bucket_size + 1).view(-1)
    pids = pid[None, :].expand(B, P)
    x = linear(feat.flatten(2).transpose(1, 2), sd, ap + ".image_proj")
    pos = F.embedding(pids, sd[ap + ".embed_image_positions.weight"])
    bias = []
    if cfg.mode == "A":
        nrd = (2 * cfg.image_bucket_size - 1) ** 2 + 3
        bucket = make_image_bucket_position(cfg.image_bucket_size, nrd)
        rp = bucket[pid][:, pid]  # double gather :118-124
        for idx in range(num_layers):
            v = F.embedding(rp, sd[f"{ap}.image_rel_pos_table_list.{idx}.weight"])  # P x P x H
            bias.append(v.permute(2, 0, 1).unsqueeze(0).expand(B, -1, -1, -1))
    return _hook(sd, ap, cfg, slot, AOut(x, masks, pos, bias), num_layers, None)


def video_adaptor(sd, gp, cfg, slot, num_layers, training=True):
    """adaptor/video_image_sequence.py:111-208: the image_resnet adaptor's backbone / image_proj / 2-D positions on
    B*F frames, frame positions (id = f + 1) added to the patch positions, padding = all-zero frames (:136-139),
    bias[(f,p),(f',p')] = frame_table[bucket(f,f')] + image_table[bucket2d(p,p')] (:176-200).  The hook
    (LayerNorms, type embedding) uses the *video* adaptor's own parameters."""
    ap, rp = gp + ".video_image_sequence", gp + ".image_resnet"
    v = slot.value.transpose(1, 2)  # B x F x 3 x H x W
    B, Fr = v.shape[:2]
    feat = resnet_backbone(v.reshape(-1, v.size(2), v.size(3), v.size(4)), sd, rp + ".embed_images", cfg.resnet_type, training)
    h, w = feat.shape[-2:]
    P = h * w
    emb = feat.reshape(feat.size(0), feat.size(1), -1).transpose(1, 2).reshape(B, Fr * P, feat.size(1))
    masks = (v.reshape(B, Fr, -1).abs().mean(dim=-1) == 0.0).unsqueeze(-1).expand(B, Fr, P).reshape(B, Fr * P)
    pid = (torch.arange(w).unsqueeze(0).expand(h, w) + torch.arange(h).unsqueeze(1) * cfg.image_bucket_size + 1).view(-1)
    ipos = F.embedding(pid[None, :].expand(B, P), sd[rp + ".embed_image_positions.weight"])  # B x P x d
    fpos = F.embedding((torch.arange(Fr) + 1)[None, :].expand(B, Fr), sd[ap + ".embed_frame_positions.weight"])  # B x F x d
    pos = (ipos.unsqueeze(1) + fpos.unsqueeze(2)).reshape(B, Fr * P, -1)
    x = linear(emb, sd, rp + ".image_proj")
    bias = []
    if cfg.mode == "A":
        nrd = (2 * cfg.image_bucket_size - 1) ** 2 + 3
        rp_img = make_image_bucket_position(cfg.image_bucket_size, nrd)[pid][:, pid]
        rp_frm = make_token_bucket_position(cfg.token_bucket_size, 1024)[:Fr, :Fr]  # make_video_bucket_position :50-60
        for idx in range(num_layers):
            vi = F.embedding(rp_img, sd[f"{rp}.image_rel_pos_table_list.{idx}.weight"]).permute(2, 0, 1)  # H x P x P
            vf = F.embedding(rp_frm, sd[f"{ap}.video_rel_pos_table_list.{idx}.weight"]).permute(2, 0, 1)  # H x F x F
            val = vf[:, :, None, :, None] + vi[:, None, :, None, :]  # H x F x P x F x P
            bias.append(val.reshape(1, val.size(0), Fr * P, Fr * P).expand(B, -1, -1, -1))
    return _hook(sd, ap, cfg, slot, AOut(x, masks, pos, bias), num_layers, None)


_DEFAULT_ADAPTOR = {  # adaptor/general.py:36-46
    TEXT: "text", IMAGE: "image_resnet", BOX: "text", AUDIO: "audio_fbank", PHONE: "text",
    VIDEO: "video_image_sequence", MOTION: "text", STRUCT: "text", CATEGORY: "text",
}
_ADAPTOR_FN = {
    "text": text_adaptor,
    "image_patch_embed": patch_embed_adaptor,
    "audio_fbank": audio_adaptor,
    "image_resnet": image_resnet_adaptor,
    "video_image_sequence": video_adaptor,
}


def general_adaptor(sd, gp, cfg: OracleConfig, slots: List[OSlot], is_src: bool):
    """adaptor/general.py:120-158 (dispatch in ModalityType order, keep slot order) + concat :245-282."""
    num_layers = cfg.enc_layers if is_src else cfg.dec_layers
    outs: List[Optional[AOut]] = [None] * len(slots)
    for mod in range(1, 10):
        for i, s in enumerate(slots):
            if s.modality == mod:
                name = s.adaptor or _DEFAULT_ADAPTOR[s.modality]
                outs[i] = _ADAPTOR_FN[name](sd, gp, cfg, s, num_layers)
    embed = torch.cat([o.embed for o in outs], dim=1)
    masks = torch.cat([o.masks for o in outs], dim=1)
    pos = torch.cat([o.pos_embed for o in outs], dim=1)
    if cfg.mode != "A":
        return embed, masks, pos, None
    B, S = pos.shape[:2]
    H = cfg.heads
    s = float(cfg.embed_dim / cfg.heads * cfg.attn_scale_factor) ** -0.5  # general.py:98
    pq = linear(pos, sd, gp + ".pos_q_linear").view(B, S, H, -1).transpose(1, 2) * s  # (s on the fp32 product in the CUDA path)
    pk = linear(pos, sd, gp + ".pos_k_linear").view(B, S, H, -1).transpose(1, 2)
    abs_bias = torch.matmul(pq, pk.transpose(2, 3))  # :223-243
    biases = []
    for idx in range(num_layers):
        b = abs_bias.clone()
        start = 0
        for o in outs:
            end = start + o.embed.shape[1]
            if o.self_attn_bias[idx] is not None:
                b[:, :, start:end, start:end] += o.self_attn_bias[idx]  # :270-280
            start = end
        biases.append(b)
    return embed, masks, pos, biases


# ------------------------------------------------------------------ encoder / decoder / model
def encoder_forward(sd, cfg: OracleConfig, slots: List[OSlot]):
    """model/transformer.py:78-156."""
    embed, masks, pos, biases = general_adaptor(sd, "encoder.adaptor", cfg, slots, True)
    has_pad = bool(masks.any())
    if has_pad:
        embed = embed * (1 - masks.unsqueeze(-1).type_as(embed))  # :109-112
    x = embed.transpose(0, 1)
    for i in range(cfg.enc_layers):
        bias = biases[i].reshape(-1, x.size(0), x.size(0)) if biases is not None else None
        x = encoder_layer(sd, f"encoder.layers.{i}", cfg, x, masks if has_pad else None, bias)
    x = layer_norm(x, sd, "encoder.layer_norm", st=True)
    return {"encoder_out": x, "encoder_padding_mask": masks, "position_embeddings": pos}


def decoder_forward(sd, cfg: OracleConfig, slots: List[OSlot], enc):
    """model/transformer.py:365-522 (+ get_cross_pos_info :280-299, buffered_future_mask :528-539)
    and the tied output projection adaptor/text.py:129-142 / base.py:131."""
    embed, masks, pos, biases = general_adaptor(sd, "decoder.adaptor", cfg, slots, False)
    B, T = embed.shape[:2]
    H = cfg.heads
    cross = None
    if cfg.mode == "A":
        s = float(cfg.embed_dim / cfg.heads * cfg.attn_scale_factor) ** -0.5
        src_pos = enc["position_embeddings"]
        S = src_pos.shape[1]
        pq = linear(pos, sd, "decoder.cross_pos_q_linear").view(B, T, H, -1).transpose(1, 2) * s
        pk = linear(src_pos, sd, "decoder.cross_pos_k_linear").view(B, S, H, -1).transpose(1, 2)
        cross = torch.matmul(pq, pk.transpose(2, 3)).reshape(-1, T, S)
    x = embed.transpose(0, 1)
    future = torch.triu(torch.full((T, T), float("-inf")), 1)
    for i in range(cfg.dec_layers):
        bias = biases[i].reshape(-1, T, T) if biases is not None else None
        x = decoder_layer(
            sd, f"decoder.layers.{i}", cfg, x, enc["encoder_out"], enc["encoder_padding_mask"], future, masks, bias, cross
        )
    x = layer_norm(x, sd, "decoder.layer_norm", st=True).transpose(0, 1)
    logits = _st(F.linear(x, sd["decoder.adaptor.embed_tokens.weight"]))
    return logits, x


def model_forward(sd: Dict[str, torch.Tensor], cfg: OracleConfig, slots: List[OSlot]):
    """model/ofa.py:165-285 (OFAEncoderDecoderExecutor.forward): returns logits [B, T, V]."""
    enc = encoder_forward(sd, cfg, [s for s in slots if s.is_src])
    logits, feats = decoder_forward(sd, cfg, [s for s in slots if not s.is_src], enc)
    return logits, {"encoder_out": enc, "last_hidden_state": feats}


def cross_entropy_sum(logits, target, pad=PAD):
    """engine/criterion/cross_entropy.py:62-67 + nll_loss :27-41: fp32 log-softmax, sum-reduced NLL
    over non-pad targets; sample_size = ntokens."""
    lprobs = F.log_softmax(logits.float(), dim=-1).view(-1, logits.size(-1))
    return F.nll_loss(lprobs, target.view(-1), ignore_index=pad, reduction="sum")


def label_smoothed_cross_entropy_sum(logits, target, epsilon, pad=PAD):
    """engine/criterion/label_smoothed_cross_entropy.py:62-92,175-191 (no constraint masks, no drop-worst): fp32
    log-softmax; over non-pad targets  loss = (1 - eps - eps_i) * nll + eps_i * smooth  with smooth = -sum_c lprobs,
    eps_i = eps / (V - 1).  Returns (loss_sum, nll_sum, ntokens)."""
    lprobs = F.log_softmax(logits.float(), dim=-1).view(-1, logits.size(-1))
    tgt = target.view(-1)
    keep = tgt != pad
    lprobs, tgt = lprobs[keep], tgt[keep]
    nll = -lprobs.gather(dim=-1, index=tgt.unsqueeze(-1)).squeeze(-1)
    smooth = -lprobs.sum(dim=-1)
    eps_i = epsilon / (lprobs.size(-1) - 1)
    loss = (1.0 - epsilon - eps_i) * nll + eps_i * smooth
    return loss.sum(), nll.sum(), loss.numel()


def label_smoothed_nll_loss(lprobs, target, epsilon, update_num, drop_worst_ratio=0.0, drop_worst_after=0, constraint_masks=None):
    """engine/criterion/label_smoothed_cross_entropy.py:62-92 in full (constraint masks :72-77, drop-worst :79-82).
    lprobs: fp32 [n, V] of the counted rows (disallowed entries -inf under constraints).  Returns (loss, nll, ntokens)."""
    nll = -lprobs.gather(dim=-1, index=target.unsqueeze(-1)).squeeze(-1)
    if constraint_masks is not None:
        smooth = -lprobs.masked_fill(~constraint_masks, 0).sum(dim=-1)
        eps_i = epsilon / (constraint_masks.sum(1) - 1 + 1e-6)
    else:
        smooth = -lprobs.sum(dim=-1)
        eps_i = epsilon / (lprobs.size(-1) - 1)
    loss = (1.0 - epsilon - eps_i) * nll + eps_i * smooth
    if drop_worst_ratio > 0 and update_num > drop_worst_after:
        loss, idx = torch.topk(loss, k=int(loss.shape[0] * (1 - drop_worst_ratio)), largest=False)
        nll = nll[idx]
    return loss.sum(), nll.sum(), loss.numel()


def constrained_criterion(logits, target, epsilon, constraint_range=None, constraint_masks=None, update_num=0, drop_worst_ratio=0.0,
                          drop_worst_after=0, pad=PAD):
    """LabelSmoothedCrossEntropyCriterion.compute_loss with constraints (label_smoothed_cross_entropy.py:147-191):
    get_constraint_masks (range: entries [4, start) and [end, V) off, AND the sample's masks), logits masked to -inf, fp32
    log-softmax, padding rows dropped, label_smoothed_nll_loss."""
    V = logits.size(-1)
    cm = constraint_masks
    if constraint_range is not None:
        rm = torch.ones(logits.shape, dtype=torch.bool)
        rm[..., 4:constraint_range[0]] = False
        rm[..., constraint_range[1]:] = False
        cm = rm if cm is None else torch.logical_and(cm, rm)
    x = logits.float()
    if cm is not None:
        x = x.masked_fill(~cm, float("-inf"))
    lprobs = F.log_softmax(x, dim=-1).view(-1, V)
    tgt = target.view(-1)
    keep = tgt != pad
    cmk = None if cm is None else cm.reshape(-1, V)[keep]
    return label_smoothed_nll_loss(lprobs[keep], tgt[keep], epsilon, update_num, drop_worst_ratio, drop_worst_after, cmk)


def make_constraint_case(seed=11, B=3, T=7, V=203, lo=50, hi=120):
    """Seeded logits (bf16 values) / targets inside the constraint range / extra trie-style masks, for the constrained criterion."""
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(B, T, V, generator=g) * 2.0).to(torch.bfloat16)
    target = torch.randint(lo, hi, (B, T), generator=g)
    target[0, -2:] = PAD
    target[2, -1] = PAD
    masks = torch.rand(B, T, V, generator=g) > 0.3
    masks.scatter_(-1, target.unsqueeze(-1), True)  # the target is always allowed
    masks[..., :4] = True
    return logits, target, masks, (lo, hi)


def make_ls_case(seed=7, rows=37, V=1003, eps=0.1):
    """Seeded logits (bf16 values) / targets with padding for the label-smoothed criterion fixtures."""
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(rows, V, generator=g) * 2.0).to(torch.bfloat16)
    target = torch.randint(4, V, (rows,), generator=g)
    target[::5] = PAD
    return logits, target, eps


# ------------------------------------------------------------------ helpers for tests / bench
def make_cfg_from_state_dict(sd, mode, **kw) -> OracleConfig:
    d = sd["encoder.adaptor.embed_tokens.weight"].shape[1]
    V = sd["encoder.adaptor.embed_tokens.weight"].shape[0]
    H = sd["encoder.layers.0.self_attn.c_attn"].shape[0]
    ffn_dim = sd["encoder.layers.0.fc1.weight"].shape[0]
    ne = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.layers."))
    nd = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.layers."))
    return OracleConfig(embed_dim=d, heads=H, ffn_dim=ffn_dim, enc_layers=ne, dec_layers=nd, vocab=V, mode=mode, **kw)


def loss_and_grads(sd, cfg, slots, target):
    """fwd + bwd of the measured path: returns (loss, logits, {name: grad})."""
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k and "version" not in k
             and not k.endswith("rp_bucket")]
    # tied embedding: one leaf shared under both names (general.py:191-221)
    leaf = {}
    params = {}
    for k in names:
        v = sd[k]
        key = v.data_ptr()
        if key not in leaf:
            leaf[key] = v.detach().requires_grad_(True)  # shares storage: no 1 GB copy per step
        params[k] = leaf[key]
    full = dict(sd)
    full.update(params)
    logits, _ = model_forward(full, cfg, slots)
    loss = cross_entropy_sum(logits, target)
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    return loss.detach(), logits.detach(), grads
