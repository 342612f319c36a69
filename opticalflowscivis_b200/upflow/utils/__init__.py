"""Module path mirror of UPFlow/utils (only what sits on the hot path: pytorch_correlation.Corr_pyTorch)."""
