"""GPU-side pieces of the image / box preprocessors at the path's entry.

`normalize_images`: ToTensor + Normalize of `DefaultImagePreprocess` (preprocessor/default/image.py:93-116) on decoded, resized
uint8 pixels [B, H, W, 3] -> [B, 3, H, W]; the PIL decode / bicubic resize stay in the data loader.
`box_to_tokens`: the `<bin>_k` quantisation of `DefaultBoxPreprocess` (preprocessor/default/box.py:101-110) on device.
No CPU fallback."""
import ctypes

import torch

from .. import _lib

IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
IMAGENET_INCEPTION_MEAN, IMAGENET_INCEPTION_STD = (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)  # the reference's default (image.py:99-104)


def normalize_images(pixels, mean=IMAGENET_INCEPTION_MEAN, std=IMAGENET_INCEPTION_STD, out_dtype=torch.float32):
    if not pixels.is_cuda or pixels.dtype != torch.uint8 or pixels.dim() != 4 or pixels.shape[-1] != 3:
        raise _lib.OfabError("normalize_images needs a CUDA uint8 tensor [B, H, W, 3] (no CPU fallback)")
    pixels = pixels.contiguous()
    B, H, W, _ = pixels.shape
    out = torch.empty((B, 3, H, W), dtype=out_dtype, device=pixels.device)
    m3, s3 = (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)
    _lib.call("ofab_image_normalize", ctypes.c_void_p(pixels.data_ptr()), B, H, W, m3, s3, ctypes.c_void_p(out.data_ptr()),
              _lib.F32 if out_dtype == torch.float32 else _lib.BF16, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return out


def box_to_tokens(coords, first_bin_id, max_image_size=512, num_bins=1000):
    """coords fp32 CUDA [..] -> int64 token ids `first_bin_id + round(x / max_image_size * (num_bins - 1))`."""
    if not coords.is_cuda:
        raise _lib.OfabError("box_to_tokens needs a CUDA tensor (no CPU fallback)")
    c = coords.to(torch.float32).contiguous()
    out = torch.empty(c.shape, dtype=torch.int64, device=c.device)
    _lib.call("ofab_box_bins", ctypes.c_void_p(c.data_ptr()), c.numel(), float(max_image_size), int(num_bins), int(first_bin_id),
              ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return out
