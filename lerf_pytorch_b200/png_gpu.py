"""Result images as PNG files, assembled on the GPU (output side).

The reference saves every result with PIL on the host (``resample/eval_lut_sr.py:667-708``); deflating a 2K x4 result
costs seconds per image there.  ``encode_png`` wraps a uint8 HWC image that is already in device memory into a complete PNG
file on the device (``lerf_png_encode_stored``: filter type 0, stored deflate blocks, Adler-32 and CRC-32 by parallel
kernels) and ``save_png`` copies the bytes to the host and writes them.  Lossless and uncompressed: the file is the image's
size plus 0.01 %.  There is no CPU fallback.
"""
import torch

from . import _lib
from .lut_interp import _stream_ptr

__all__ = ["encode_png", "save_png", "png_bytes"]


def png_bytes(H, W, channels):
    """Size in bytes of the PNG file ``encode_png`` produces for an H x W x channels image."""
    n = _lib.lib().lerf_png_stored_bytes(int(H), int(W), int(channels))
    if n < 0:
        raise ValueError("PNG of a %dx%dx%d image: bad shape or more than 2^31-1 bytes of image data" % (H, W, channels))
    return int(n)


def encode_png(img, out=None):
    """img: uint8 CUDA tensor [H,W] or [H,W,C] (C = 1, 2, 3, 4).  Returns a uint8 CUDA tensor holding the PNG file (a view of
    ``out`` if given: a uint8 CUDA tensor of at least ``png_bytes`` rounded up to 4 bytes, 16-byte aligned)."""
    if not isinstance(img, torch.Tensor) or img.dtype != torch.uint8 or not img.is_cuda or img.dim() not in (2, 3):
        raise ValueError("encode_png needs a uint8 CUDA tensor [H,W] or [H,W,C]")
    H, W = int(img.shape[0]), int(img.shape[1])
    C = 1 if img.dim() == 2 else int(img.shape[2])
    n = png_bytes(H, W, C)
    cap = (n + 3) & ~3
    dev = img.device
    if out is None:
        out = torch.empty(cap, dtype=torch.uint8, device=dev)
    elif out.dtype != torch.uint8 or not out.is_cuda or out.numel() < cap or not out.is_contiguous():
        raise ValueError("encode_png: `out` must be a contiguous uint8 CUDA tensor of at least %d bytes" % cap)
    scratch = torch.empty(4, dtype=torch.int64, device=dev)
    src = img.contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().lerf_png_encode_stored(src.data_ptr(), H, W, C, out.data_ptr(), out.numel(), scratch.data_ptr(),
                                                     _stream_ptr(dev)))
    return out[:n]


def save_png(img, path):
    """Encode on the device, copy the file to the host, write it to ``path``.  Returns the number of bytes written."""
    data = encode_png(img).cpu().numpy()
    with open(path, "wb") as f:
        f.write(data.tobytes())
    return int(data.size)
