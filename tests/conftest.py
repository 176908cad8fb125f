"""Shared fixtures.  `-m "not gpu"` runs on any machine (oracle, host logic, ABI surface); `-m gpu` needs a B200."""
from __future__ import annotations

import gzip
import hashlib
import importlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
if REPO not in sys.path:
    sys.path.insert(0, REPO)

SCENES = ("hw15_scene2", "hw09_scene5", "hw11_scene8", "hw12_scene4")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def quantise(rgb: np.ndarray) -> np.ndarray:
    """uint8(255.999 * clamp(c, 0, 1)), product in double - io/image/ppm.hpp:17-19."""
    c = np.clip(rgb.astype(np.float32), np.float32(0), np.float32(1)).astype(np.float64)
    return (255.999 * c).astype(np.uint8)


def psnr8(a: np.ndarray, b: np.ndarray) -> float:
    d = quantise(a).astype(np.float64) - quantise(b).astype(np.float64)
    mse = float((d * d).mean())
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)


def check_textures_png(got8: np.ndarray, published8: np.ndarray) -> None:
    """an 8-bit config-4 frame (hw12/scene4, spp 1) against the decoded outputs/textures.png (README.md:64-65): EVERY pixel - the
    albedo, edges and checker quadrants and, since the fixture's texels come from a restatement of the reference's JPEG decoder
    arithmetic (tests/helpers/jpeg_stb.py), the bitmap quadrant too"""
    assert got8.shape == published8.shape == (1080, 1920, 3)
    assert np.array_equal(got8, published8), f"{int((got8 != published8).any(axis=2).sum())} pixels differ from outputs/textures.png"


@pytest.fixture(scope="session")
def golden() -> dict:
    with open(os.path.join(HERE, "golden", "golden.json")) as fh:
        return json.load(fh)


_scene_cache: dict[str, bytes] = {}


def scene_bytes(name: str) -> bytes:
    if name not in _scene_cache:
        with gzip.open(os.path.join(HERE, "golden", "scenes", name + ".rtsc.gz"), "rb") as fh:
            _scene_cache[name] = fh.read()
    return _scene_cache[name]


def resized(data: bytes, width: int, height: int) -> bytes:
    """RTSC bytes with another image size (width, height live at byte offsets 20 and 24)."""
    b = bytearray(data)
    b[20:24] = int(width).to_bytes(4, "little")
    b[24:28] = int(height).to_bytes(4, "little")
    return bytes(b)


@pytest.fixture(scope="session")
def rt():
    """The product binding.  Builds the library when it is missing (nvcc cross-compiles without a GPU)."""
    lib = os.path.join(REPO, "simd-raytracer_b200", "librt_b200.so")
    if not os.path.exists(lib):
        sys.path.insert(0, os.path.join(REPO, "simd-raytracer_b200"))
        import build as _build  # noqa: PLC0415
        _build.build()
    return importlib.import_module("simd-raytracer_b200")


@pytest.fixture(scope="session")
def oracle_mod():
    from tests.helpers import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def records():
    def load(name: str):
        return np.load(os.path.join(HERE, "golden", f"records_{name}.npz"))
    return load
