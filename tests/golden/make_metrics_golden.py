"""Pins oracle/metrics_ref.py against the reference's own error.py (imported unmodified from /root/reference, cv2 required)
and writes tests/golden/metrics.npz.  Run in the build container only:  python tests/golden/make_metrics_golden.py"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import metrics_ref as mr                                   # noqa: E402

# error.py imports plotting helpers it does not need for the metrics: stub them (SURVEY.md Appendix C recipe)
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "utils"):
    m = types.ModuleType(name)
    m.use = lambda *a, **k: None
    m.visualize_ind = m.visualize_series = lambda *a, **k: None
    sys.modules.setdefault(name, m)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
spec = importlib.util.spec_from_file_location("ref_error", "/root/reference/error.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.RandomState(1234)
cases = {}
log = []
for tag, shape in (("a", (160, 224)), ("b", (64, 96)), ("c", (48, 40, 1)), ("d", (40, 56, 3))):
    base = rng.randint(0, 256, size=shape).astype(np.float32)
    smooth = base.copy()
    for _ in range(3):       # some spatial structure: SSIM of pure noise is degenerate
        smooth = 0.25 * (np.roll(smooth, 1, 0) + np.roll(smooth, -1, 0) + np.roll(smooth, 1, 1) + np.roll(smooth, -1, 1))
    img1 = np.clip(smooth * 2.0 - 100.0, 0, 255).astype(np.float32)
    img2 = np.clip(img1 + rng.normal(0, 6.0, size=shape), 0, 255).astype(np.float32)
    p_ref, s_ref = ref.calculate_psnr(img1, img2), ref.calculate_ssim(img1, img2)
    p_or, s_or = mr.calculate_psnr(img1, img2), mr.calculate_ssim(img1, img2)
    assert abs(p_ref - p_or) <= 1e-12 * abs(p_ref), (tag, p_ref, p_or)
    assert abs(s_ref - s_or) <= 1e-10, (tag, s_ref, s_or)
    cases[f"{tag}_img1"], cases[f"{tag}_img2"] = img1, img2
    cases[f"{tag}_psnr"], cases[f"{tag}_ssim"] = np.float64(p_ref), np.float64(s_ref)
    log.append(f"metrics {tag} {shape}: psnr {p_ref:.6f} (oracle diff {abs(p_ref - p_or):.1e}), ssim {s_ref:.9f} (oracle diff {abs(s_ref - s_or):.1e})")
assert ref.calculate_psnr(cases["a_img1"], cases["a_img1"]) == float("inf") == mr.calculate_psnr(cases["a_img1"], cases["a_img1"])
k_ref = __import__("cv2").getGaussianKernel(11, 1.5)[:, 0]
assert np.abs(k_ref - mr.gaussian_kernel(11, 1.5)).max() <= 1e-16, np.abs(k_ref - mr.gaussian_kernel(11, 1.5)).max()
log.append(f"gaussian kernel vs cv2.getGaussianKernel(11, 1.5): max diff {np.abs(k_ref - mr.gaussian_kernel(11, 1.5)).max():.1e}")
np.savez_compressed(os.path.join(HERE, "metrics.npz"), **cases)
with open(os.path.join(HERE, "PINNING.txt"), "a") as f:
    f.write("\n".join(log) + "\n")
print("\n".join(log))
