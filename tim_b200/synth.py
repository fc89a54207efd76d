"""Deterministic synthetic weights and inputs (no RNG library state involved).

A counter-based splitmix64 hash gives every tensor a reproducible stream keyed by
(seed, tensor name), so the golden-vector generator (tools/make_golden.py, which runs the
real reference), the oracle tests, the GPU parity tests and bench.py all see bit-identical
fp32 inputs without shipping the tensors themselves.

Input distributions follow SURVEY.md §8(d): features ~ N(0,1); feature times on a regular grid
(datasets/sliding_window.py:363-404 normalises window times to [0,1]); query times
start ~ U(0,0.9), len ~ U(0.01,0.3).
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

from .config import TIMConfig, state_dict_spec

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _stream(seed: int, name: str, n: int, lane: int = 0) -> np.ndarray:
    """n uniform doubles in [0,1) for (seed, name, lane)."""
    key = (np.uint64(zlib.crc32(name.encode())) << np.uint64(32)) ^ np.uint64(seed & 0xFFFFFFFF)
    key = _splitmix64(np.array([key], dtype=np.uint64))[0] ^ np.uint64(lane * 0x632BE59BD9B4E019 & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95) + key
        z = _splitmix64(ctr)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform(seed: int, name: str, shape, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    n = int(np.prod(shape)) if len(shape) else 1
    u = _stream(seed, name, n)
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def normal(seed: int, name: str, shape, std: float = 1.0) -> np.ndarray:
    n = int(np.prod(shape)) if len(shape) else 1
    u1 = _stream(seed, name, n, lane=1)
    u2 = _stream(seed, name, n, lane=2)
    z = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
    return (std * z).astype(np.float32).reshape(shape)


def synth_state_dict(cfg: TIMConfig, seed: int = 0, style: str = "trained") -> "OrderedDict[str, np.ndarray]":
    """fp32 numpy state_dict with the reference's key names/shapes.

    style="init":    magnitudes of the reference's default initialisation.
    style="trained": LayerNorm gains/biases moved off 1/0, non-zero biases everywhere, larger
                     attention logits and visible CLS / modality parameters, so parity is not
                     only exercised at init statistics.
    """
    trained = style == "trained"
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for key, shape in state_dict_spec(cfg).items():
        leaf = key.rsplit(".", 1)[-1]
        is_norm = (".norm1." in key or ".norm2." in key or key.startswith("time_mlp.6.")
                   or "_embedder.3." in key)
        if is_norm:
            if leaf == "weight":
                w = 1.0 + uniform(seed, key, shape, -0.25, 0.25) if trained else np.ones(shape, np.float32)
            else:
                w = uniform(seed, key, shape, -0.15, 0.15) if trained else np.zeros(shape, np.float32)
        elif key.endswith("_cls") or key.endswith("modality_encoding"):
            w = normal(seed, key, shape, std=0.3 if trained else 0.01)
        elif leaf in ("weight", "in_proj_weight"):
            fan_in = shape[-1]
            gain = 1.0
            if trained and leaf == "in_proj_weight":
                gain = 1.6
            a = gain / np.sqrt(fan_in)
            w = uniform(seed, key, shape, -a, a)
        else:  # biases
            fan_in = 16.0
            a = 0.1 if trained else 0.02
            w = uniform(seed, key, shape, -a, a)
        sd[key] = np.ascontiguousarray(w, dtype=np.float32)
    return sd


def synth_inputs(cfg: TIMConfig, B: int, Qv: int, Qa: int, seed: int = 1234,
                 shared_queries: bool = False) -> Dict[str, np.ndarray]:
    """Synthetic clip batch: {'vis','aud','times'} (+ query split sizes).

    times is [B, T, 2] with T = F_tot + Qv + Qa in the reference's order
    (feature times for vis, then aud, then visual queries, then audio queries;
    recognition/scripts/test.py:95-117, detection/.../tim.py:345-378).

    shared_queries=True reproduces detection inference, where one query set
    (model.inference_queries) is repeated for every clip and for both modalities
    (detection/.../tim.py:348,364): clip 0's visual queries are broadcast.
    """
    F = cfg.num_feats
    out: Dict[str, np.ndarray] = {}
    if cfg.has_visual_input:
        out["vis"] = normal(seed, "vis", (B, F, cfg.visual_input_dim))
    if cfg.has_audio_input:
        out["aud"] = normal(seed + 1, "aud", (B, F, cfg.audio_input_dim))
    grid = np.arange(F, dtype=np.float32) / np.float32(F)
    ft = np.stack([grid, grid + np.float32(1.0 / F)], axis=-1)          # [F, 2]
    parts = []
    if cfg.has_visual_input:
        parts.append(np.broadcast_to(ft, (B, F, 2)))
    if cfg.has_audio_input:
        parts.append(np.broadcast_to(ft, (B, F, 2)))
    nqv = Qv if "visual" in cfg.data_modality else 0
    nqa = Qa if "audio" in cfg.data_modality else 0
    for k, (name, nq) in enumerate((("qv", nqv), ("qa", nqa))):
        if nq:
            st = uniform(seed + 2 + k, name + "_start", (B, nq), 0.0, 0.9)
            ln = uniform(seed + 4 + k, name + "_len", (B, nq), 0.01, 0.3)
            q = np.stack([st, st + ln], axis=-1)
            if shared_queries:
                if k == 1 and nqv:
                    if nqa != nqv:
                        raise ValueError("shared_queries needs Qa == Qv")
                    q = parts[-1]
                else:
                    q = np.broadcast_to(q[0:1], q.shape)
            parts.append(q)
    out["times"] = np.ascontiguousarray(np.concatenate(parts, axis=1), dtype=np.float32)
    return out


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """||a-b||_2 / ||b||_2 in float64 — the parity metric of BASELINE.md §4."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = float(np.linalg.norm(b))
    return float(np.linalg.norm(a - b)) / (den if den > 0 else 1.0)
