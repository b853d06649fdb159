"""AudioFbankAdaptor, source branch (ofasys/adaptor/audio.py:190-325) with Conv2dSubsampling4
(ofasys/module/subsample.py:11-63).

conv1 (1->d, 3x3 s2)+ReLU is a direct kernel; conv2 (d->d, 3x3 s2: ~90 % of the adaptor's FLOPs) is
im2col + the tcgen05 GEMM; activations stay channel-last [B, T, F, C], so the reference's NCHW weight
and its Linear(C*F', d) feature order (c*F' + f) are mapped by a small [O, A, B] -> [O, B, A] weight
permute each step instead of permuting activations.

The TTS-side sub-modules of the reference adaptor (prenet / postnet / feat_proj / eos_proj / mask_emb)
are parameters only (checkpoint compatibility); they are not on the ASR path.
"""
from dataclasses import dataclass

import torch
import torch.nn as nn

from .. import ops
from ..configure import register_config
from ..module import Embedding
from .base import AdaptorOutput, BaseAdaptor, BaseAdaptorConfig
from .text import make_token_bucket_position

DEFAULT_MAX_WAV_POSITIONS = 4096


def make_audio_bucket_1d(bucket_size, max_position=DEFAULT_MAX_WAV_POSITIONS):
    """bucket id as a function of (i - j) in [-(max_position-1), max_position-1]: the reference's
    [4096, 4096] int64 buffer (audio.py:50-60, 134 MB on each of encoder and decoder) depends only on
    the difference, so one row of it is kept; elementwise float ops are identical, hence bit-exact."""
    import math
    rel = torch.arange(-(max_position - 1), max_position, dtype=torch.long)
    sign = torch.sign(rel)
    mid = bucket_size // 2
    abs_pos = torch.where((rel < mid) & (rel > -mid), mid - 1, torch.abs(rel))
    log_pos = torch.ceil(torch.log(abs_pos / mid) / math.log((max_position - 1) / mid) * (mid - 1)) + mid
    log_pos = log_pos.int()
    bucket = torch.where(abs_pos.le(mid), rel, log_pos * sign).long()
    return bucket + bucket_size - 1


class Conv2dSubsampling4(nn.Module):
    """Parameter container with the reference's names: conv.0, conv.2, out.0 (subsample.py:27-33)."""

    def __init__(self, idim: int, odim: int):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(1, odim, 3, 2), nn.ReLU(), nn.Conv2d(odim, odim, 3, 2), nn.ReLU())
        self.out = nn.Sequential(nn.Linear(odim * (((idim - 1) // 2 - 1) // 2), odim))
        self.subsampling_rate = 4
        self.right_context = 6

    def get_out_seq_lens_tensor(self, in_seq_lens_tensor):
        out = in_seq_lens_tensor.clone()
        for _ in range(2):
            out = ((out.float() - 1) / 2 + 1).floor().long()  # quirk 7: one frame more than the conv yields
        return out

    def forward(self, x, x_length):
        """x: [B, L, idim] -> bf16 [B, T', odim]."""
        B, L, F = x.shape
        c1, c2, lin = self.conv[0], self.conv[2], self.out[0]
        C = c1.out_channels
        y1 = ops.conv1_relu(x, c1.weight, c1.bias)  # [B, H1, W1, C]
        H1, W1 = y1.shape[1:3]
        H2, W2 = (H1 - 3) // 2 + 1, (W1 - 3) // 2 + 1
        cols = ops.im2col_3x3s2(y1)  # [B*H2*W2, 9C]
        w2 = ops.transpose_last2(c2.weight.reshape(C, C, 9)).reshape(C, 9 * C)  # [co, (kh,kw), ci]
        y2 = ops.relu(ops.linear(cols, w2, c2.bias))  # [B*H2*W2, C] == [B, H2, W2*C]
        wl = ops.transpose_last2(lin.weight.reshape(lin.out_features, C, W2)).reshape(lin.out_features, W2 * C)
        out = ops.linear(y2.view(B * H2, W2 * C), wl, lin.bias).view(B, H2, lin.out_features)
        return out, self.get_out_seq_lens_tensor(x_length)


@dataclass
class AudioFbankAdaptorConfig(BaseAdaptorConfig):
    output_frame_dim: int = 80
    n_frames_per_step: int = 1
    is_transformer_layers: bool = False
    prenet_layers: int = 2
    prenet_dim: int = 256
    postnet_conv_dim: int = 512
    postnet_conv_kernel_size: int = 5
    postnet_layers: int = 5
    use_mask: bool = False


@register_config("ofasys.adaptor", "audio_fbank", AudioFbankAdaptorConfig)
class AudioFbankAdaptor(BaseAdaptor):
    def __init__(self, embed_tokens, dictionary, is_src, general_adaptor, cfg: AudioFbankAdaptorConfig):
        super().__init__(embed_tokens, dictionary, is_src, general_adaptor, cfg)
        assert not cfg.is_transformer_layers, "is_transformer_layers is broken in the reference (quirk 9) and unsupported"
        self.audio_bucket_size = cfg.max_position
        self.out_dim = cfg.output_frame_dim * cfg.n_frames_per_step
        self.subsample = Conv2dSubsampling4(self.out_dim, cfg.embed_dim)
        self.embed_audio_positions = Embedding(cfg.max_position, cfg.embed_dim)
        audio_num_rel_dis = 2 * self.audio_bucket_size - 1
        n_tables = 1 if cfg.share_attn_bias else self.num_layers
        self.audio_rel_pos_table_list = nn.ModuleList(
            [Embedding(audio_num_rel_dis, cfg.num_attention_heads, zero_init=True) for _ in range(n_tables)]
        )
        # non-persistent: derived data, 64 KB instead of the reference's 134 MB `audio_rp_bucket` buffer
        self.register_buffer("audio_rp_bucket_1d", make_audio_bucket_1d(self.audio_bucket_size), persistent=False)
        self._idx_cache = {}

    def rel_idx(self, T):
        k = (T, self.audio_rp_bucket_1d.device)
        if k not in self._idx_cache:
            i = torch.arange(T, device=self.audio_rp_bucket_1d.device)
            rel = i[:, None] - i[None, :] + (DEFAULT_MAX_WAV_POSITIONS - 1)
            self._idx_cache[k] = self.audio_rp_bucket_1d[rel].to(torch.int32).contiguous()
        return self._idx_cache[k]

    def forward(self, slot, **kwargs) -> AdaptorOutput:
        if not slot.is_src:
            raise NotImplementedError("audio target (TTS) branch is outside the hot path")
        fbank = slot.value["fbank"]
        lens = slot.value["fbank_lengths"]
        feat, out_lens = self.subsample(fbank, lens)
        B, T = feat.shape[:2]
        # audio.py:307-310 without the per-sample host loop: frames >= reported length are padding
        masks = torch.arange(T, device=feat.device)[None, :] >= out_lens[:, None]
        embed, pos = self.hook(slot, self.embed_audio_positions.weight, dense=feat, zero_mask=masks)
        out = AdaptorOutput(embed, masks, None if pos is None else pos.expand(B, -1, -1), None)
        if self.cfg.use_self_attn_bias:
            out.rel_idx = self.rel_idx(T)
            out.rel_tables = [t.weight for t in self.audio_rel_pos_table_list]
        return out

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        state_dict.pop(prefix + "audio_rp_bucket", None)  # reference checkpoints carry the 134 MB table
        for k in [k for k in state_dict if k.startswith(prefix) and k[len(prefix):].split(".")[0] in
                  ("prenet", "postnet", "feat_proj", "eos_proj", "mask_emb")]:
            state_dict.pop(k)  # TTS-only parameters of the reference adaptor
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
