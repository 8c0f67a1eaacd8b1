"""Quality metrics of the reference's LUT eval scripts, restated (host side, numpy / torch CPU).

* ``rgb2ycbcr``  <- common/utils.py:46-76  ``_rgb2ycbcr``
* ``psnr``       <- common/utils.py:138-151 ``PSNR`` (float32 arithmetic, border shave)
* ``ssim``       <- common/utils.py:177-203 ``cal_ssim`` (11x11 Gaussian window, sigma 1.5, 'valid')
* ``mpsnr``      <- common/utils.py:168-175 ``mPSNR`` (masked PSNR of the warp benchmark, torch float32)

They exist so the eval-script adapters (eval_lut_sr.py / eval_lut_warp.py in this package) print the reference's
table without importing the reference.  Same operations in the same order and precision; scipy does the window
convolution exactly like the reference (its requirements list scipy too).
"""
import numpy as np

_T = np.array([[0.256788235294118, 0.504129411764706, 0.097905882352941],
               [-0.148223529411765, -0.290992156862745, 0.439215686274510],
               [0.439215686274510, -0.367788235294118, -0.071427450980392]])
_O = (16, 128, 128)


def rgb2ycbcr(img, max_val=255):
    """[H,W,3] RGB -> [H,W,3] YCbCr (float64), ITU-R BT.601 'studio swing' like MATLAB's rgb2ycbcr."""
    off = np.array(_O, dtype=np.float64) / (255.0 if max_val == 1 else 1.0)
    t = np.reshape(img, (img.shape[0] * img.shape[1], img.shape[2]))
    t = np.dot(t, np.transpose(_T))
    t[:, 0] += off[0]
    t[:, 1] += off[1]
    t[:, 2] += off[2]
    return np.reshape(t, [img.shape[0], img.shape[1], img.shape[2]])


def psnr(y_true, y_pred, shave_border=4):
    """Inputs 0..255, 2-D.  float32 like the reference."""
    target = np.array(y_true, dtype=np.float32)
    ref = np.array(y_pred, dtype=np.float32)
    diff = ref - target
    if shave_border > 0:
        diff = diff[shave_border:-shave_border, shave_border:-shave_border]
    rmse = np.sqrt(np.mean(np.power(diff, 2)))
    return 20 * np.log10(255. / rmse)


def _gaussian_kernel(ksize=11, sigma=1.5):
    """cv2.getGaussianKernel(ksize, sigma): exp(-(i-(ksize-1)/2)^2 / (2 sigma^2)), normalised, column vector."""
    i = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-(i * i) / (2.0 * sigma * sigma))
    return (k / k.sum()).reshape(-1, 1)


def ssim(img1, img2):
    from scipy import signal
    K = [0.01, 0.03]
    L = 255
    kx = _gaussian_kernel(11, 1.5)
    window = kx * kx.T
    C1 = (K[0] * L) ** 2
    C2 = (K[1] * L) ** 2
    img1 = np.float64(img1)
    img2 = np.float64(img2)
    mu1 = signal.convolve2d(img1, window, 'valid')
    mu2 = signal.convolve2d(img2, window, 'valid')
    mu1_sq = mu1 * mu1
    mu2_sq = mu2 * mu2
    mu1_mu2 = mu1 * mu2
    sigma1_sq = signal.convolve2d(img1 * img1, window, 'valid') - mu1_sq
    sigma2_sq = signal.convolve2d(img2 * img2, window, 'valid') - mu2_sq
    sigma12 = signal.convolve2d(img1 * img2, window, 'valid') - mu1_mu2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return np.mean(ssim_map)


def mpsnr(sr, hr, mask, rgb_range=255):
    """Masked PSNR (torch float32 CPU like the reference); sr, hr, mask: arrays of one shape, mask in {0,1}."""
    import torch
    sr, hr, mask = torch.Tensor(np.asarray(sr)), torch.Tensor(np.asarray(hr)), torch.Tensor(np.asarray(mask))
    diff = mask * (sr - hr) / rgb_range
    gain = mask.nelement() / mask.sum()
    mse = gain.item() * diff.pow(2).mean()
    return float(-10 * torch.log10(mse))


def psnr_y_ssim(img_gt, img_out, scale_h, scale_w):
    """The metric block at the end of eltr._worker (resample/eval_lut_sr.py:735-744): crop to the common size,
    Y channel, PSNR with shave = max(int(scale)), SSIM."""
    if img_gt.shape != img_out.shape:
        predH, predW, _ = img_out.shape
        img_gt = img_gt[:predH, :predW, :]
        gtH, gtW, _ = img_gt.shape
        img_out = img_out[:gtH, :gtW, :]
    y_gt, y_out = rgb2ycbcr(img_gt)[:, :, 0], rgb2ycbcr(img_out)[:, :, 0]
    return [psnr(y_gt, y_out, max(int(scale_h), int(scale_w))), ssim(y_gt, y_out)]
