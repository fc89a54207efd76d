#!/usr/bin/env python
"""Mints tests/golden/grads.npz: torch.autograd gradients of the UNMODIFIED reference forward (float64, CPU) - the pin of
oracle/tim_oracle_bwd.py, groundwork for the training leg.

    python tools/make_golden_grads.py            (build container only; one subprocess per variant, like tools/make_golden.py)

For each case (configs, seeded weights and inputs of tools/make_golden.py) the reference runs with gradients enabled and dropout 0
(model.eval()), the scalar L = sum_k <output_k, cotangent_k> with seeded normal cotangents for every output is back-propagated,
and for every parameter (and the two feature inputs) the file keeps a fingerprint of the gradient: its L2 norm, its sum, and its
values at up to 512 seeded flat indices - enough to pin a restatement without committing 60 MB of gradients for the real widths.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.make_golden import CASES, INPUT_SEED, WEIGHT_SEED, build_reference   # noqa: E402

GRAD_CASES = ["recog_cfg1", "recog_av_small", "recog_av_novn", "recog_av_qa0", "recog_visual", "recog_audio", "recog_hd192",
              "det_visual", "det_visual_vn", "det_av", "det_audio"]
N_SAMPLES = 512


def cotangent(case: str, key: str, shape) -> np.ndarray:
    """Seeded N(0,1) cotangent of output `key` (float64); tests regenerate it with the same call."""
    rng = np.random.default_rng(zlib.crc32(f"{case}/{key}".encode()))
    return rng.standard_normal(tuple(shape))


def sample_indices(case: str, key: str, numel: int) -> np.ndarray:
    rng = np.random.default_rng(zlib.crc32(f"idx/{case}/{key}".encode()))
    return np.sort(rng.choice(numel, size=min(N_SAMPLES, numel), replace=False))


def fingerprint(case: str, key: str, grad: np.ndarray):
    flat = np.asarray(grad, np.float64).reshape(-1)
    idx = sample_indices(case, key, flat.size)
    return np.array([np.sqrt((flat * flat).sum()), flat.sum()]), flat[idx]


def run_variant(variant: str):
    import torch
    from tim_b200.config import TIMConfig
    from tim_b200.synth import synth_inputs, synth_state_dict
    torch.set_num_threads(8)
    blob = {}
    for name in GRAD_CASES:
        kw, B, Qv, Qa = CASES[name]
        cfg = TIMConfig(**kw)
        if cfg.variant != variant:
            continue
        model = build_reference(cfg).double()
        sd = synth_state_dict(cfg, WEIGHT_SEED, "trained")
        model.load_state_dict({k: torch.from_numpy(v).double() for k, v in sd.items()}, strict=True)
        det = cfg.variant == "detection"
        inp = synth_inputs(cfg, B, Qv, Qa, INPUT_SEED, shared_queries=det)
        vis = torch.from_numpy(inp["vis"]).double().requires_grad_(True) if "vis" in inp else None
        aud = torch.from_numpy(inp["aud"]).double().requires_grad_(True) if "aud" in inp else None
        times = torch.from_numpy(inp["times"]).double()
        if not det:
            te = model(times, "time_mlp")
            (verb, noun, action, audio), feats = model([vis, aud], "encoder", te, Qv, Qa)
            reg_v = reg_a = None
        else:
            nq = max(Qv, Qa)
            model.inference_queries = times[0:1, cfg.F_tot:cfg.F_tot + nq].clone()
            model.num_queries = nq
            res = model([vis, aud], "encoder", times[:, :cfg.F_tot].clone(), None, False)
            (verb, noun, action, audio), (reg_v, reg_a), feats = res[0]
        loss = 0.0
        for k, v in (("verb", verb), ("noun", noun), ("action", action), ("audio", audio), ("reg_v", reg_v), ("reg_a", reg_a),
                     ("feats", feats)):
            if v is not None:
                loss = loss + (v * torch.from_numpy(cotangent(name, k, v.shape))).sum()
                blob[f"{name}/out_shape/{k}"] = np.array(v.shape)
        loss.backward()
        grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        if vis is not None and vis.grad is not None:
            grads["input.vis"] = vis.grad
        if aud is not None and aud.grad is not None:
            grads["input.aud"] = aud.grad
        for k, gr in grads.items():
            st, vals = fingerprint(name, k, gr.numpy())
            blob[f"{name}/stat/{k}"], blob[f"{name}/vals/{k}"] = st, vals
        blob[f"{name}/keys"] = np.array(sorted(grads))
        nograd = sorted(k for k, p in model.named_parameters() if p.grad is None)
        blob[f"{name}/no_grad"] = np.array(nograd)
        print(f"[grads] {name}: {len(grads)} gradient tensors, loss {float(loss):.6f}, without gradient: {nograd}", flush=True)
    return blob


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--variant":
        blob = run_variant(sys.argv[2])
        np.savez_compressed(sys.argv[3], **blob)
        return
    merged = {}
    for variant in ("recognition", "detection"):
        tmp = os.path.join(ROOT, "tests", "golden", f"_grads_{variant}.npz")
        r = subprocess.run([sys.executable, __file__, "--variant", variant, tmp], capture_output=True, text=True)
        sys.stdout.write(r.stdout)
        if r.returncode:
            sys.stderr.write(r.stderr)
            raise SystemExit(r.returncode)
        with np.load(tmp) as z:
            merged.update({k: z[k] for k in z.files})
        os.remove(tmp)
    merged["cases"] = np.array(GRAD_CASES)
    out = os.path.join(ROOT, "tests", "golden", "grads.npz")
    np.savez_compressed(out, **merged)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
