"""Recursive / ratio-targeted interpolation drivers around `Model.inference`.

Mirrors the two loops of `Flow-2D/inference_img.py:64-97` (same in `Flow-3D/inference_img.py`): `--exp k` doubles the
sequence k times (2^k - 1 new members between the two inputs), `--ratio r` bisects towards time r.  The reference scripts
call an HD model whose `inference` returns the middle frame; `select` picks it out of this package's tuples (2-D
`merged[2]`, 3-D `merged`).  Everything stays on the device; no arithmetic happens here.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch


def _middle(out):
    return out[0] if torch.is_tensor(out[0]) else out[0][2]


def interpolate_recursive(model, img0: torch.Tensor, img1: torch.Tensor, exp: int = 4,
                          select: Optional[Callable] = None) -> List[torch.Tensor]:
    """inference_img.py:88-97 — returns the 2^exp + 1 members [img0, ..., img1]."""
    select = select or _middle
    img_list = [img0, img1]
    for _ in range(exp):
        tmp = []
        for j in range(len(img_list) - 1):
            mid = select(model.inference(img_list[j], img_list[j + 1]))
            tmp.append(img_list[j])
            tmp.append(mid)
        tmp.append(img1)
        img_list = tmp
    return img_list


def interpolate_ratio(model, img0: torch.Tensor, img1: torch.Tensor, ratio: float, rthreshold: float = 0.02,
                      rmaxcycles: int = 8, select: Optional[Callable] = None) -> List[torch.Tensor]:
    """inference_img.py:64-87 — bisection towards time `ratio` in (0, 1); returns [img0, middle, img1]."""
    select = select or _middle
    img0_ratio, img1_ratio = 0.0, 1.0
    if ratio <= img0_ratio + rthreshold / 2:
        middle = img0
    elif ratio >= img1_ratio - rthreshold / 2:
        middle = img1
    else:
        tmp_img0, tmp_img1 = img0, img1
        middle = None
        for _ in range(rmaxcycles):
            middle = select(model.inference(tmp_img0, tmp_img1))
            middle_ratio = (img0_ratio + img1_ratio) / 2
            if ratio - (rthreshold / 2) <= middle_ratio <= ratio + (rthreshold / 2):
                break
            if ratio > middle_ratio:
                tmp_img0, img0_ratio = middle, middle_ratio
            else:
                tmp_img1, img1_ratio = middle, middle_ratio
    return [img0, middle, img1]
