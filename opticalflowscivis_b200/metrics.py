"""Device-side evaluation metrics with the reference's names and meaning.

Mirrors `error.py:27-103` (`calculate_psnr`, `ssim`, `calculate_ssim`, `calculate_metrics`; numpy float64 on [0,255] images
there) and the validation PSNR of `Flow-3D/train.py:385-388` / `Flow-2D/train.py:477-480` (`-10*log10(mean((gt-pred)^2))`
on [0,1] tensors).  Inputs are float32 CUDA tensors; the sums run in float64 inside libofsv (`ofsv_sq_err_f64`,
`ofsv_ssim2d_f64`) in a fixed order, only the scalar results come back to the host.
"""
from __future__ import annotations

import math
import statistics
from typing import Sequence, Tuple

import torch

from . import ops


def calculate_psnr(img1: torch.Tensor, img2: torch.Tensor) -> float:
    """error.py:27-34 — img1, img2 on the [0,255] scale, any shape; `inf` for identical inputs."""
    if img1.shape != img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    sse = float(ops.sq_err_sums(img1.reshape(1, -1), img2.reshape(1, -1))[0])
    mse = sse / img1.numel()
    if mse == 0:
        return float("inf")
    return 20 * math.log10(255.0 / math.sqrt(mse))


def psnr_per_sample(pred: torch.Tensor, gt: torch.Tensor) -> list:
    """Flow-3D/train.py:385-388 — one `-10*log10(mean((gt[j]-pred[j])^2))` per batch member, [0,1] tensors (N, ...)."""
    sse = ops.sq_err_sums(gt, pred).cpu()
    count = pred.numel() // pred.shape[0]
    return [-10 * math.log10(float(s) / count) if float(s) > 0 else float("inf") for s in sse]


def ssim(img1: torch.Tensor, img2: torch.Tensor) -> float:
    """error.py:36-56 — (H, W) images, or (H, W, C) filtered per channel like cv2.filter2D; mean of the SSIM map."""
    if img1.dim() == 2:
        return float(ops.ssim2d_means(img1, img2)[()] if img1.dim() == 2 else 0.0)
    if img1.dim() == 3:
        a, b = img1.permute(2, 0, 1).contiguous(), img2.permute(2, 0, 1).contiguous()
        return float(ops.ssim2d_means(a, b).mean())
    raise ValueError("Wrong input image dimensions.")


def calculate_ssim(img1: torch.Tensor, img2: torch.Tensor) -> float:
    """error.py:58-76 — same dispatch and error behaviour as the reference (note its 3-channel branch averages three
    identical whole-image calls, reproduced here as one)."""
    if not img1.shape == img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if img1.dim() == 2:
        return ssim(img1, img2)
    elif img1.dim() == 3:
        if img1.shape[2] == 3:
            return ssim(img1, img2)
        elif img1.shape[2] == 1:
            return ssim(img1[:, :, 0], img2[:, :, 0])
        return None            # the reference falls through without a value for other channel counts
    else:
        raise ValueError("Wrong input image dimensions.")


def calculate_metrics(original_data: Sequence[torch.Tensor], interpol_data: Sequence[torch.Tensor], factor: int) -> Tuple[float, float]:
    """error.py:78-103 — mean PSNR / SSIM over the INTERPOLATED members of a sequence (indices i % factor != 0)."""
    psnr_i, ssim_i = [], []
    for i in range(min(len(original_data), len(interpol_data))):
        if i % factor != 0:
            psnr_i.append(calculate_psnr(original_data[i], interpol_data[i]))
            ssim_i.append(calculate_ssim(original_data[i], interpol_data[i]))
    return statistics.mean(psnr_i), statistics.mean(ssim_i)
