"""Image statistics of the GI parity bound: PSNR on 8-bit frames, PSNR after an n x n box filter,
channel means.  Test infrastructure only."""
from __future__ import annotations

import numpy as np


def psnr_u8(a: np.ndarray, b: np.ndarray) -> float:
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float((d * d).mean())
    return 99.0 if mse == 0 else float(10.0 * np.log10(255.0 * 255.0 / mse))


def box(a: np.ndarray, n: int) -> np.ndarray:
    """mean over non-overlapping n x n pixel blocks (the frame is cropped to a multiple of n)"""
    h, w = (a.shape[0] // n) * n, (a.shape[1] // n) * n
    return a[:h, :w].astype(np.float64).reshape(h // n, n, w // n, n, -1).mean(axis=(1, 3))


def psnr_box(a: np.ndarray, b: np.ndarray, n: int) -> float:
    d = box(a, n) - box(b, n)
    mse = float((d * d).mean())
    return 99.0 if mse == 0 else float(10.0 * np.log10(255.0 * 255.0 / mse))


def channel_means(a: np.ndarray) -> np.ndarray:
    return a.reshape(-1, a.shape[-1]).astype(np.float64).mean(0)
