"""VideoImageSequenceAdaptor (ofasys/adaptor/video_image_sequence.py:63-208): the image_resnet adaptor's backbone,
projection and 2-D patch positions applied to every frame of a clip [B, 3, F, H, W]; frame positions (id = f + 1)
are added to the patch positions; all-zero frames are padding; the relative-position bias of token pair
((f,p),(f',p')) is frame_table[bucket(f,f')] + image_table[bucket2d(p,p')].

The bias stays structured: one int32 [F*P, F*P] id map into a per-layer product table
`frame_table[used_f][:, None] + image_table[used_p][None, :]` (only the ids a clip of this shape can produce), which
the attention kernel gathers from shared memory -- the reference materialises [B, H, F*P, F*P] per layer."""
from dataclasses import dataclass

import torch
import torch.nn as nn

from .. import ops
from ..configure import ConfigStore, register_config
from ..module import Embedding
from .base import AdaptorOutput, BaseAdaptor, BaseAdaptorConfig
from .image_resnet import ImageResnetAdaptor
from .text import make_token_bucket_position


def make_video_bucket_position(bucket_size, max_position=8192):
    """frame index pair -> bucket (video_image_sequence.py:50-60; the text adaptor's function)."""
    return make_token_bucket_position(bucket_size, max_position)


@dataclass
class VideoImageSequenceAdaptorConfig(BaseAdaptorConfig):
    token_bucket_size: int = 256


@register_config("ofasys.adaptor", "video_image_sequence", VideoImageSequenceAdaptorConfig)
class VideoImageSequenceAdaptor(BaseAdaptor):
    def __init__(self, embed_tokens, dictionary, is_src, general_adaptor, cfg: VideoImageSequenceAdaptorConfig):
        super().__init__(embed_tokens, dictionary, is_src, general_adaptor, cfg)
        self.embed_frame_positions = Embedding(1024 + 1, cfg.embed_dim, zero_init=True)
        video_num_rel_dis = 2 * cfg.token_bucket_size - 1
        n_tables = 1 if cfg.share_attn_bias else self.num_layers
        self.video_rel_pos_table_list = nn.ModuleList(
            [Embedding(video_num_rel_dis, cfg.num_attention_heads, zero_init=True) for _ in range(n_tables)]
        )
        self.register_buffer("video_rp_bucket", make_video_bucket_position(cfg.token_bucket_size, 1024))
        ga = self.general_adaptor
        if "image_resnet" not in ga.name2adaptor and is_src:  # :84-96: the frames run through the image_resnet adaptor
            icfg = getattr(ga.cfg.adaptor, "image_resnet")
            icfg.parse_from_model_cfg(ga.cfg)
            ga.name2adaptor["image_resnet"] = ConfigStore().get("ofasys.adaptor", "image_resnet").target(
                embed_tokens, dictionary, is_src, ga, icfg)
            setattr(ga, "image_resnet", ga.name2adaptor["image_resnet"])
        self._cache = {}

    def get_image_resnet_adaptor(self) -> ImageResnetAdaptor:
        a = self.general_adaptor.name2adaptor["image_resnet"]
        assert a is not None
        return a

    def _index(self, frames, h, w, device):
        """(used frame ids, used patch ids, int32 [F*P, F*P] ids into the [n_f * n_p] product table)"""
        k = (frames, h, w, device)
        if k not in self._cache:
            _, rel_p = self.get_image_resnet_adaptor().position_ids(h, w, device)  # int32 [P, P]
            rel_f = self.video_rp_bucket.to(device)[:frames, :frames]
            used_f, inv_f = torch.unique(rel_f, return_inverse=True)
            used_p, inv_p = torch.unique(rel_p.long(), return_inverse=True)
            P = h * w
            idx = inv_f[:, None, :, None] * used_p.numel() + inv_p[None, :, None, :]  # F x P x F x P
            self._cache[k] = (used_f, used_p, idx.reshape(frames * P, frames * P).to(torch.int32).contiguous())
        return self._cache[k]

    def forward(self, slot, **kwargs) -> AdaptorOutput:
        clip = slot.value
        B, C, Fr = clip.shape[:3]
        ira = self.get_image_resnet_adaptor()
        frames, zero = ops.video_frames(clip)  # bf16 [B*F, 3, H, W], bool [B, F]
        x, h, w = ira.features(frames)  # [B*F, P, d]
        P = h * w
        pid, _ = ira.position_ids(h, w, clip.device)
        ipos = ira.embed_image_positions.weight.index_select(0, pid)  # [P, d]
        fpos = self.embed_frame_positions.weight[1:Fr + 1]  # ids f + 1
        pos_rows = (fpos.unsqueeze(1) + ipos.unsqueeze(0)).reshape(Fr * P, -1)
        embed, pos = self.hook(slot, pos_rows, dense=x.view(B, Fr * P, -1))
        masks = zero.unsqueeze(-1).expand(B, Fr, P).reshape(B, Fr * P)
        out = AdaptorOutput(embed, masks, None if pos is None else pos.expand(B, -1, -1), None)
        if self.cfg.use_self_attn_bias:
            used_f, used_p, idx = self._index(Fr, h, w, clip.device)
            out.rel_idx = idx
            out.rel_tables = [
                (tf.weight.index_select(0, used_f).float().unsqueeze(1) + ti.weight.index_select(0, used_p).float().unsqueeze(0)).reshape(-1, tf.weight.shape[1])
                for tf, ti in zip(self.video_rel_pos_table_list, ira.image_rel_pos_table_list)
            ]
        return out
