"""The eval scripts' quality metrics on the GPU ("GPU PSNR-Y/SSIM/mPSNR so metrics don't become the
new bottleneck"): the super-resolved image never has to leave the device for a metrics-only run.

Same definitions as ``metrics.py`` (the host restatement of common/utils.py:46-76, :138-151, :168-175, :177-203 of the
reference), evaluated with torch CUDA ops in float64 (float32 where the reference uses float32).  Summation order differs
from numpy's, so results agree with the host versions to ~1e-6 dB / 1e-9 SSIM, far below the two / four decimals the
reference prints.  These are reporting utilities, not part of the hot path.
"""
import torch

from .metrics import _T, _O, _gaussian_kernel


def _y_channel(img_u8_hwc):
    """[H,W,3] uint8 CUDA -> [H,W] float64 luma (rgb2ycbcr(...)[:, :, 0])."""
    t = torch.tensor(_T[0], dtype=torch.float64, device=img_u8_hwc.device)
    return img_u8_hwc.to(torch.float64) @ t + float(_O[0])


def psnr(y_true, y_pred, shave_border=4):
    """common/utils.py:138-151 on [H,W] CUDA tensors (float32 arithmetic like the reference)."""
    diff = y_pred.to(torch.float32) - y_true.to(torch.float32)
    if shave_border > 0:
        diff = diff[shave_border:-shave_border, shave_border:-shave_border]
    rmse = torch.sqrt(torch.mean(diff * diff))
    return float(20 * torch.log10(255.0 / rmse))


def ssim(img1, img2):
    """common/utils.py:177-203 (11x11 Gaussian window, sigma 1.5, 'valid'), float64; the window is separable, so the five
    2-D convolutions are done as row and column passes."""
    k = torch.tensor(_gaussian_kernel(11, 1.5)[:, 0], dtype=torch.float64, device=img1.device)
    x = torch.stack([img1, img2, img1 * img1, img2 * img2, img1 * img2]).to(torch.float64).unsqueeze(1)  # [5,1,H,W]
    x = torch.nn.functional.conv2d(x, k.view(1, 1, 1, -1))
    x = torch.nn.functional.conv2d(x, k.view(1, 1, -1, 1))[:, 0]
    mu1, mu2 = x[0], x[1]
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq, sigma2_sq, sigma12 = x[2] - mu1_sq, x[3] - mu2_sq, x[4] - mu1_mu2
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return float(ssim_map.mean())


def mpsnr(sr, hr, mask, rgb_range=255):
    """common/utils.py:168-175 on CUDA tensors of one shape (mask in {0,1}), float32 like the reference."""
    sr, hr, mask = sr.to(torch.float32), hr.to(torch.float32), mask.to(torch.float32)
    diff = mask * (sr - hr) / rgb_range
    gain = mask.numel() / float(mask.sum())
    return float(-10 * torch.log10(gain * (diff * diff).mean()))


def psnr_y_ssim(img_gt, img_out, scale_h, scale_w):
    """The metric block of eltr._worker (resample/eval_lut_sr.py:735-744) on uint8 [H,W,3] CUDA tensors."""
    if img_gt.shape != img_out.shape:
        predH, predW, _ = img_out.shape
        img_gt = img_gt[:predH, :predW, :]
        gtH, gtW, _ = img_gt.shape
        img_out = img_out[:gtH, :gtW, :]
    y_gt, y_out = _y_channel(img_gt), _y_channel(img_out)
    return [psnr(y_gt, y_out, max(int(scale_h), int(scale_w))), ssim(y_gt, y_out)]
