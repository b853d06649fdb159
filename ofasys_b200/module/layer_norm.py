"""LayerNorm factory + module (ofasys/module/layer_norm.py:27-32): same constructor, same
`weight`/`bias` parameter names; forward runs the sm_100a kernels of csrc/ln.cu."""
import torch
import torch.nn as nn

from .. import ops


class FusedLayerNorm(nn.LayerNorm):
    def forward(self, x, out_dtype=torch.bfloat16):
        assert self.elementwise_affine, "ofasys_b200 LayerNorm needs elementwise_affine"
        return ops.layer_norm(x, self.weight, self.bias, self.eps, gelu=False, out_dtype=out_dtype)


def LayerNorm(normalized_shape, eps=1e-5, elementwise_affine=True, export=False):
    return FusedLayerNorm(normalized_shape, eps, elementwise_affine)
