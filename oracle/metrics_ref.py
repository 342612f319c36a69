"""TEST INFRASTRUCTURE — numpy float64 restatement of the reference's evaluation metrics.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path (opticalflowscivis_b200)
never does.  Pinned against the reference itself (error.py imported unmodified, with cv2) by tests/golden/make_metrics_golden.py;
the fixture tests/golden/metrics.npz replays that comparison without the reference or cv2.
"""
from __future__ import annotations

import math

import numpy as np


def calculate_psnr(img1, img2):
    """error.py:27-34."""
    img1 = np.asarray(img1).astype(np.float64)
    img2 = np.asarray(img2).astype(np.float64)
    mse = np.mean((img1 - img2) ** 2)
    if mse == 0:
        return float("inf")
    return 20 * math.log10(255.0 / math.sqrt(mse))


def gaussian_kernel(ksize: int = 11, sigma: float = 1.5):
    """cv2.getGaussianKernel(ksize, sigma) for sigma > 0 (OpenCV imgproc/smooth: exp(-(i - (n-1)/2)^2 / (2 sigma^2)), then
    multiplied by the reciprocal of the sum), float64."""
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-0.5 / (sigma * sigma) * x * x)
    return k * (1.0 / k.sum())


def _filter_valid(img, window):
    """cv2.filter2D(img, -1, window)[5:-5, 5:-5] == 'valid' correlation (the border mode never reaches the kept region)."""
    kh, kw = window.shape
    H, W = img.shape[:2]
    out = np.zeros((H - kh + 1, W - kw + 1) + img.shape[2:], dtype=np.float64)
    for dy in range(kh):
        for dx in range(kw):
            out += window[dy, dx] * img[dy:dy + H - kh + 1, dx:dx + W - kw + 1]
    return out


def ssim(img1, img2):
    """error.py:36-56."""
    C1 = (0.01 * 255) ** 2
    C2 = (0.03 * 255) ** 2
    img1 = np.asarray(img1).astype(np.float64)
    img2 = np.asarray(img2).astype(np.float64)
    kernel = gaussian_kernel(11, 1.5)
    window = np.outer(kernel, kernel)
    mu1 = _filter_valid(img1, window)
    mu2 = _filter_valid(img2, window)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    sigma1_sq = _filter_valid(img1 ** 2, window) - mu1_sq
    sigma2_sq = _filter_valid(img2 ** 2, window) - mu2_sq
    sigma12 = _filter_valid(img1 * img2, window) - mu1_mu2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def calculate_ssim(img1, img2):
    """error.py:58-76."""
    if not img1.shape == img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if img1.ndim == 2:
        return ssim(img1, img2)
    elif img1.ndim == 3:
        if img1.shape[2] == 3:
            return np.array([ssim(img1, img2) for _ in range(3)]).mean()
        elif img1.shape[2] == 1:
            return ssim(np.squeeze(img1), np.squeeze(img2))
    else:
        raise ValueError("Wrong input image dimensions.")


def psnr_train(pred, gt):
    """Flow-3D/train.py:385 — -10*log10(mean((gt - pred)^2)), float64 here."""
    d = np.asarray(gt, dtype=np.float64) - np.asarray(pred, dtype=np.float64)
    return -10 * math.log10(np.mean(d * d))
