"""Shared helpers for the test-suite: golden fixtures and synthetic streams."""
from __future__ import annotations

import hashlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

GOLDEN_NAMES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def load_golden_stream(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("kind", "mode", "dec_sha256", "coefs_sha256", "ranks_sha256"):
        g[k] = str(g[k])
    for k in ("stream_id", "n_samples", "block_size", "n_chan", "rate_hz"):
        g[k] = int(g[k])
    for k in ("param", "avg_complexity"):
        g[k] = float(g[k])
    return g


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_pcm(g: dict) -> np.ndarray:
    from ulc_codec_b200 import corpus
    return corpus.generate(g["stream_id"], g["n_samples"], g["kind"])


def peak_rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| relative to the peak of b (SURVEY.md 8a parity note 5)."""
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(1e-30, np.abs(b).max()))


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.sqrt((d * d).sum() / max(1e-30, (b.astype(np.float64) ** 2).sum())))
