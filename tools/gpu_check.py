#!/usr/bin/env python
"""Diagnostic sweep on a B200 (not a test, never stops at the first failure): single-kernel checks through the
C-ABI test hooks, then every golden case and the named configs against the numpy oracle. Prints one line per check."""
import ctypes as C
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tim_b200 import _lib                                  # noqa: E402
from tim_b200.config import TIMConfig, named_config        # noqa: E402
from tim_b200.plugin import TIMEngine                      # noqa: E402
from tim_b200.synth import synth_inputs, synth_state_dict, rel_l2   # noqa: E402

DT = {"fp32": 0, "bf16": 1, "fp16": 2}
TORCH_DT = {"bf16": torch.bfloat16, "fp16": torch.float16}
dev = torch.device("cuda", 0)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def check_linear(lib):
    g = torch.Generator(device="cpu").manual_seed(0)
    shapes = [(128, 256, 64), (128, 256, 128), (300, 1024, 1024), (77, 97, 128), (1000, 3072, 1024), (256, 64, 2304),
              (200, 512, 512), (5000, 1024, 2048), (129, 300, 72), (20000, 2048, 1024)]
    for (M, N, K) in shapes:
        A = torch.randn(M, K, generator=g).to(dev)
        W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev)
        bias = torch.randn(N, generator=g).to(dev)
        resid = torch.randn(M, N, generator=g).to(dev)
        for dt in ("fp32", "fp16", "bf16"):
            for act, use_res in ((0, False), (2, True), (1, False)):
                out = torch.full((M, N), float("nan"), device=dev)
                try:
                    r = lib.tim_test_linear(DT[dt], ptr(A), ptr(W), ptr(bias), ptr(resid if use_res else None), ptr(out),
                                            M, N, K, act, C.c_void_p(0))
                    if r != 0:
                        print(f"linear {dt} M={M} N={N} K={K} act={act}: status {r} {lib.tim_last_error(None)}")
                        continue
                    torch.cuda.synchronize()
                except Exception as e:  # noqa
                    print(f"linear {dt} M={M} N={N} K={K}: EXC {e}")
                    continue
                if dt == "fp32":
                    a, w = A.double(), W.double()
                else:
                    a, w = A.to(TORCH_DT[dt]).double(), W.to(TORCH_DT[dt]).double()
                ref = a @ w.T + bias.double()
                if act == 1:
                    ref = torch.relu(ref)
                elif act == 2:
                    ref = torch.nn.functional.gelu(ref)
                if use_res:
                    ref = ref + resid.double()
                err = rel_l2(out.cpu().numpy(), ref.cpu().numpy())
                nan = int(torch.isnan(out).sum())
                print(f"linear {dt:5s} M={M:6d} N={N:5d} K={K:5d} act={act} res={int(use_res)}: rel_l2={err:.3e} nan={nan}", flush=True)


def attn_ref(qkv, B, Ft, Qt, H, hd):
    """dense masked attention in fp64 on the two-stream layout (q already scaled, log2 domain)."""
    E = H * hd
    S = Ft + Qt
    out = torch.empty(qkv.shape[0], E, dtype=torch.float64)
    mask = torch.ones(S, S, dtype=torch.bool)
    mask[:, :Ft] = False
    mask.fill_diagonal_(False)
    for b in range(B):
        rows = torch.cat([torch.arange(b * Ft, (b + 1) * Ft), B * Ft + torch.arange(b * Qt, (b + 1) * Qt)])
        x = qkv[rows].double()
        q, k, v = x[:, :E], x[:, E:2 * E], x[:, 2 * E:]
        q = q.view(S, H, hd).transpose(0, 1); k = k.view(S, H, hd).transpose(0, 1); v = v.view(S, H, hd).transpose(0, 1)
        sc = (q @ k.transpose(1, 2)) * math.log(2.0)          # log2 domain -> natural
        sc = sc.masked_fill(mask[None], float("-inf"))
        p = torch.softmax(sc, -1)
        o = (p @ v).transpose(0, 1).reshape(S, E)
        out[rows] = o
    return out


def check_attention(lib):
    g = torch.Generator(device="cpu").manual_seed(1)
    cases = [(2, 12, 7, 2, 16), (3, 100, 100, 2, 128), (2, 128, 200, 1, 192), (2, 100, 0, 2, 64), (1, 50, 300, 2, 32),
             (2, 100, 75, 8, 128), (1, 1, 5, 1, 16)]
    for (B, Ft, Qt, H, hd) in cases:
        E = H * hd
        M = B * (Ft + Qt)
        qkv = torch.randn(M, 3 * E, generator=g)
        qkv[:, :E] *= hd ** -0.5 * 1.4426950408889634 * 2.0    # sharper logits than unit scale
        for dt in ("fp32", "fp16", "bf16"):
            x = qkv if dt == "fp32" else qkv.to(TORCH_DT[dt]).float()
            out = torch.full((M, E), float("nan"), device=dev)
            xin = x.to(dev).contiguous()
            r = lib.tim_test_attention(DT[dt], ptr(xin), ptr(out), B, Ft, Qt, H, hd, C.c_void_p(0))
            if r != 0:
                print(f"attn {dt} B={B} Ft={Ft} Qt={Qt} H={H} hd={hd}: status {r} {lib.tim_last_error(None)}")
                continue
            torch.cuda.synchronize()
            ref = attn_ref(x, B, Ft, Qt, H, hd)
            err = rel_l2(out.cpu().numpy(), ref.numpy())
            print(f"attn {dt:5s} B={B} Ft={Ft:3d} Qt={Qt:3d} H={H} hd={hd:3d}: rel_l2={err:.3e} nan={int(torch.isnan(out).sum())}", flush=True)


def run_engine(cfg, sd, inp, Qv, Qa, dt, host=False):
    eng = TIMEngine(cfg, 0, dt)
    eng.load_state_dict(sd)
    vis = torch.from_numpy(inp["vis"]) if "vis" in inp else None
    aud = torch.from_numpy(inp["aud"]) if "aud" in inp else None
    times = torch.from_numpy(inp["times"])
    if host:
        o, up, down = eng.forward_host(vis.pin_memory() if vis is not None else None,
                                       aud.pin_memory() if aud is not None else None, times.pin_memory(), Qv, Qa,
                                       clips_per_chunk=max(1, times.shape[0] // 3))
        res = {k: (v.numpy().copy() if v is not None else None) for k, v in o.items()}
    else:
        te = eng.time_mlp(times.to(dev))
        o = eng.encoder(vis.to(dev) if vis is not None else None, aud.to(dev) if aud is not None else None, te, Qv, Qa)
        torch.cuda.synchronize()
        res = {k: (v.cpu().numpy() if v is not None else None) for k, v in o.items()}
        res["time_encodings"] = te.cpu().numpy()
    n = eng.launch_count
    eng.close()
    return res, n


def check_golden():
    gold_dir = os.path.join(ROOT, "tests", "golden")
    man = json.load(open(os.path.join(gold_dir, "manifest.json")))["cases"]
    from tests.test_oracle_golden import load_case
    for name in sorted(man):
        cfg, sd, inp, gold, c = load_case(name)
        for dt in ("fp32", "fp16", "bf16"):
            for host in (False, True):
                try:
                    res, n = run_engine(cfg, sd, inp, c["Qv"], c["Qa"], dt, host)
                except Exception as e:  # noqa
                    print(f"golden {name} {dt} host={host}: EXC {e}", flush=True)
                    continue
                errs = {k: rel_l2(res[k], v) for k, v in gold.items() if res.get(k) is not None}
                miss = [k for k in gold if res.get(k) is None and not (host and k == "time_encodings")]
                worst = max(errs.values()) if errs else float("nan")
                print(f"golden {name:18s} {dt:5s} host={int(host)} launches={n:3d} worst={worst:.3e} missing={miss} "
                      + " ".join(f"{k}={e:.1e}" for k, e in errs.items()), flush=True)


def check_named():
    from oracle.tim_oracle import TIMOracle
    for name, B in (("cfg2", 3), ("cfg3", 2), ("cfg4", 1)):
        cfg, Qv, Qa = named_config(name)
        sd = synth_state_dict(cfg, 0, "trained")
        inp = synth_inputs(cfg, B, Qv, Qa, 1234 + 10 * int(name[-1]), shared_queries=cfg.variant == "detection")
        t0 = time.time()
        ref = TIMOracle(cfg, sd, np.float32).forward(inp.get("vis"), inp.get("aud"), inp["times"], Qv, Qa, clip_chunk=1)
        t1 = time.time()
        for dt in ("fp32", "fp16", "bf16"):
            try:
                res, n = run_engine(cfg, sd, inp, Qv, Qa, dt)
            except Exception as e:  # noqa
                print(f"named {name} {dt}: EXC {e}", flush=True)
                continue
            errs = {k: rel_l2(res[k], v) for k, v in ref.items() if v is not None and res.get(k) is not None}
            print(f"named {name} B={B} {dt:5s} launches={n} oracle_s={t1 - t0:.1f} worst={max(errs.values()):.3e} "
                  + " ".join(f"{k}={e:.1e}" for k, e in errs.items()), flush=True)


if __name__ == "__main__":
    lib = _lib.load()
    print(torch.cuda.get_device_name(0), flush=True)
    what = sys.argv[1:] or ["linear", "attention", "golden", "named"]
    if "linear" in what:
        check_linear(lib)
    if "attention" in what:
        check_attention(lib)
    if "golden" in what:
        check_golden()
    if "named" in what:
        check_named()
