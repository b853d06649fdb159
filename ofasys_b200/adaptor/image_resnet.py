"""ImageResnetAdaptor (ofasys/adaptor/image_resnet.py:69-202): ResNet C1-C4 features -> Linear(1024, d),
2-D learned positions (id = w + h*bucket_size + 1) and the 2-D relative-position bucket table."""
from dataclasses import dataclass

import torch
import torch.nn as nn

from .. import ops
from ..configure import register_config
from ..module import Embedding, Linear
from ..module.resnet import resnet50_backbone, resnet101_backbone, resnet152_backbone
from .base import AdaptorOutput, BaseAdaptor, BaseAdaptorConfig


def make_image_bucket_position(bucket_size, num_relative_distance):
    """(patch id, patch id) -> relative-position bucket (image_resnet.py:25-40); integer, bit-exact."""
    coords_h = torch.arange(bucket_size)
    coords_w = torch.arange(bucket_size)
    coords = torch.stack(torch.meshgrid([coords_h, coords_w], indexing="ij"))
    coords_flatten = torch.flatten(coords, 1)
    rel = coords_flatten[:, :, None] - coords_flatten[:, None, :]
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


@dataclass
class ImageResnetAdaptorConfig(BaseAdaptorConfig):
    resnet_type: str = "resnet152"
    resnet_drop_path_rate: float = 0.0
    sync_bn: bool = False
    freeze_resnet: bool = False
    image_bucket_size: int = 42
    pretrained_ckpt_path: str = ""


@register_config("ofasys.adaptor", "image_resnet", ImageResnetAdaptorConfig)
class ImageResnetAdaptor(BaseAdaptor):
    def __init__(self, embed_tokens, dictionary, is_src, general_adaptor, cfg: ImageResnetAdaptorConfig):
        super().__init__(embed_tokens, dictionary, is_src, general_adaptor, cfg)
        assert not cfg.sync_bn and cfg.resnet_drop_path_rate == 0.0, \
            "sync_bn / resnet drop-path are off in every BASELINE config and not implemented"
        self.embed_image_positions = Embedding(cfg.image_bucket_size**2 + 1, cfg.embed_dim)
        backbone = {"resnet50": resnet50_backbone, "resnet101": resnet101_backbone, "resnet152": resnet152_backbone}[cfg.resnet_type]
        self.embed_images = backbone()
        self.image_proj = Linear(1024, cfg.embed_dim)
        image_num_rel_dis = (2 * cfg.image_bucket_size - 1) * (2 * cfg.image_bucket_size - 1) + 3
        n_tables = 1 if cfg.share_attn_bias else self.num_layers
        self.image_rel_pos_table_list = nn.ModuleList(
            [Embedding(image_num_rel_dis, cfg.num_attention_heads, zero_init=True) for _ in range(n_tables)]
        )
        self.register_buffer("image_rp_bucket", make_image_bucket_position(cfg.image_bucket_size, image_num_rel_dis))
        self._cache = {}

    def train(self, mode=True):
        """freeze_resnet (image_resnet.py:107-114): the backbone's BatchNorms stay in eval mode with frozen affine
        parameters -- the fine-tuning recipe of docs/source/howto/train.rst:66-69 -- while the convolutions keep training."""
        super().train(mode)
        if self.cfg.freeze_resnet:
            for m in self.embed_images.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
                    m.weight.requires_grad = False
                    m.bias.requires_grad = False
        return self

    def position_ids(self, h, w, device):
        k = (h, w, device)
        if k not in self._cache:
            pid = (torch.arange(w, device=device).unsqueeze(0).expand(h, w)
                   + torch.arange(h, device=device).unsqueeze(1) * self.cfg.image_bucket_size + 1).reshape(-1)
            rel = self.image_rp_bucket.to(device)[pid][:, pid].to(torch.int32).contiguous()  # double gather (image_resnet.py:118-124)
            self._cache[k] = (pid, rel)
        return self._cache[k]

    def features(self, images):
        """images [N, 3, H, W] -> (projected embeddings bf16 [N, h*w, d], h, w)"""
        feat = self.embed_images(images)
        N, h, w, C = feat.shape
        x = ops.linear(feat.view(N, h * w, C), self.image_proj.weight, self.image_proj.bias)
        return x, h, w

    def forward(self, slot, **kwargs) -> AdaptorOutput:
        images = slot.value
        B = images.shape[0]
        x, h, w = self.features(images)
        P = h * w
        pid, rel = self.position_ids(h, w, images.device)
        pos_rows = self.embed_image_positions.weight.index_select(0, pid)  # [P, d] rows in sequence order
        embed, pos = self.hook(slot, pos_rows, dense=x)
        masks = torch.zeros((B, P), dtype=torch.bool, device=images.device)
        out = AdaptorOutput(embed, masks, None if pos is None else pos.expand(B, -1, -1), None)
        if self.cfg.use_self_attn_bias:
            out.rel_idx = rel
            out.rel_tables = [t.weight for t in self.image_rel_pos_table_list]
        return out
