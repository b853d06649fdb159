"""ImagePatchEmbedAdaptor (ofasys/adaptor/image_patch_embed.py:38-80): Conv2d(3, d, k=s=patch) as
im2col + tcgen05 GEMM, CLS token, learned positions.  Needs Mode B (use_self_attn_bias=False,
entangle_position_embedding=True): the reference raises otherwise."""
from dataclasses import dataclass

import torch
import torch.nn as nn

from .. import ops
from ..configure import register_config
from ..module import Embedding
from .base import AdaptorOutput, BaseAdaptor, BaseAdaptorConfig


@dataclass
class ImagePatchEmbedAdaptorConfig(BaseAdaptorConfig):
    image_size_width: int = 224
    image_size_height: int = 224
    patch_size_width: int = 14
    patch_size_height: int = 14
    embed_dim: int = 768
    add_cls_token: bool = True


@register_config("ofasys.adaptor", "image_patch_embed", ImagePatchEmbedAdaptorConfig)
class ImagePatchEmbedAdaptor(BaseAdaptor):
    def __init__(self, embed_tokens, dictionary, is_src, general_adaptor, cfg: ImagePatchEmbedAdaptorConfig):
        super().__init__(embed_tokens, dictionary, is_src, general_adaptor, cfg)
        assert cfg.patch_size_height == cfg.patch_size_width, "square patches only"
        self.image_size = (cfg.image_size_height, cfg.image_size_width)
        self.patch_size = (cfg.patch_size_height, cfg.patch_size_width)
        self.num_patches = (self.image_size[1] // self.patch_size[1]) * (self.image_size[0] // self.patch_size[0])
        self.embed_image_positions = Embedding(self.num_patches + 1 if cfg.add_cls_token else self.num_patches, cfg.embed_dim)
        if cfg.add_cls_token:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, cfg.embed_dim))
        self.proj = nn.Conv2d(3, cfg.embed_dim, kernel_size=self.patch_size, stride=self.patch_size)

    def forward(self, slot, **kwargs) -> AdaptorOutput:
        image = slot.value
        B, C, H, W = image.shape
        assert (H, W) == self.image_size, f"Input image size ({H}*{W}) doesn't match model ({self.image_size[0]}*{self.image_size[1]})."
        if self.cfg.use_self_attn_bias or not self.cfg.entangle_position_embedding:
            raise NotImplementedError(
                "image_patch_embed returns no self_attn_bias: it needs use_self_attn_bias=False and "
                "entangle_position_embedding=True, as in the reference (base.py:183-189)")
        p = self.patch_size[0]
        d = self.cfg.embed_dim
        k = C * p * p
        kp = (k + 7) // 8 * 8
        cols = ops.patch_im2col(image, p, kp)
        w = self.proj.weight.reshape(d, k)
        if kp != k:
            w = torch.nn.functional.pad(w, (0, kp - k))  # zero columns: 16-byte TMA row stride
        patches = ops.linear(cols, w, self.proj.bias).view(B, self.num_patches, d)
        T = self.num_patches + (1 if self.cfg.add_cls_token else 0)
        embed, _ = self.hook(slot, self.embed_image_positions.weight, dense=patches, cls=self.cls_token if self.cfg.add_cls_token else None)
        masks = torch.zeros((B, T), dtype=torch.bool, device=image.device)
        pos = self.embed_image_positions.weight[:T].unsqueeze(0).expand(B, -1, -1)
        return AdaptorOutput(embed, masks, pos, None)
