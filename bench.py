#!/usr/bin/env python
"""bench.py — clips x queries / sec of the TIM forward (BASELINE.json metric) on N B200s of one node.

A "step" is one forward (time_mlp + encoder: embedders, token assembly, L encoder layers, heads) over one batch of
synthetic clips per GPU. Workload at N=1 is BASELINE.json configs[1] (cfg2: EPIC-100 recognition, 6 layers, d=512,
8 heads, 50 vis + 50 aud feature tokens, 25 + 25 interval queries = 100 query tokens).

    python bench.py --gpus 1 --steps K --warmup W          # our arm (CUDA path through the C ABI)
    python bench.py --impl reference ...                   # the reference algorithm on the host cores (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...      # one rank per GPU, clips sharded, no data-path collective

One JSON line on stdout (rank 0). `value` is device-resident throughput (inputs already in HBM); `e2e` is the same
metric through the plugin call on pinned HOST buffers with H2D / D2H copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tim_b200.config import named_config   # noqa: E402

METRIC = "clips_x_queries_per_sec"
UNIT = "clips*queries/s"
CPU_SAMPLE_CLIPS = {"cfg1": 64, "cfg2": 24, "cfg3": 6, "cfg4": 2}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d["bf16_tflops_sustained"]), "tflops_burst": float(d["bf16_tflops"]),
                "hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.2 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return None
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].isdigit() else None,
                "power_w_max": max(float(r[2]) for r in rows if r[2].replace(".", "").isdigit()) if rows else None,
                "samples": len(rows), "reasons": reasons}


def cpu_port(cfg, Qv, Qa, clips, steps, warmup, seed=1234):
    """Times the reference algorithm on the host cores: oracle/tim_oracle_torch.py, which issues op for op the PyTorch CPU calls the
    reference's modules make (dense S x S masked attention through F.multi_head_attention_forward, the materialised [B*H, S, S]
    mask, nn.Linear / LayerNorm / GELU functionals) and is bit-identical to the reference's output. The reference's own Python
    package cannot travel to the GPU box; the torch-independent numpy oracle that judges parity computes the same numbers but is
    2.9x slower than the reference on the same cores, so it is NOT what is timed here."""
    import torch
    from oracle.tim_oracle_torch import TIMOracleTorch
    from tim_b200.synth import synth_inputs, synth_state_dict
    sd = synth_state_dict(cfg, 0, "trained")
    inp = synth_inputs(cfg, clips, Qv, Qa, seed, shared_queries=cfg.variant == "detection")
    o = TIMOracleTorch(cfg, sd)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    # all host threads, whatever the launcher put in the environment: torchrun exports OMP_NUM_THREADS=1 to its workers when
    # nproc-per-node > 1, which would pin the CPU arm to one thread at N > 1
    prev = torch.get_num_threads()
    torch.set_num_threads(cores)
    ts = []
    try:
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            o.forward(inp.get("vis"), inp.get("aud"), inp["times"], Qv, Qa)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
        used = torch.get_num_threads()
    finally:
        torch.set_num_threads(prev)
    t = sum(ts) / len(ts)
    return {"value": clips * (Qv + Qa) / t, "unit": UNIT, "cores": min(cores, used), "kind": "port",
            "sample": f"{clips} clips/step x {steps} steps of the same workload, fp32, the reference's PyTorch CPU calls restated op for op "
                      f"(torch {torch.__version__}, {min(cores, used)} intra-op threads), {t * 1e3:.0f} ms/step", "ms_per_step": t * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4"])
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU per step (0 = workload default)")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--chunk", type=int, default=0, help="clips per H2D/compute/D2H chunk of the e2e path (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg, Qv, Qa = named_config(args.workload)
    wl_desc = {"cfg1": "recog L=1 d=512 25+25 feats 5+5 queries", "cfg2": "EPIC-100 recognition L=6 d=512 H=8, 50 vis + 50 aud tokens, 25+25 interval queries (100 query tokens, S=200)",
               "cfg3": "Perception-Test L=6 d=768, 64+64 tokens, 200+200 queries (S=528)", "cfg4": "detection dense queries L=6 d=512, 50+50 tokens, 2048 interval queries (S=2148)"}[args.workload]

    # ------------------------------------------------------------------ reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        clips = args.clips or CPU_SAMPLE_CLIPS[args.workload]
        steps = max(1, min(args.steps, 10))
        cb = cpu_port(cfg, Qv, Qa, clips, steps, min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 1), "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.workload}: {wl_desc}", "clips_per_step": clips,
                           "note": "reference algorithm (dense masked S x S attention) restated with the PyTorch CPU calls the reference makes, "
                                   "bit-identical to its output, run on the host cores; the Python reference itself is not present on the GPU box"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from tim_b200.plugin import TIMEngine
    from tim_b200.synth import rel_l2, synth_inputs, synth_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the TIM forward has no CPU path (use --impl reference for the CPU arm)")
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner does) goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    from tim_b200.dist import bind_to_gpu_numa_node
    numa_node = bind_to_gpu_numa_node(local_rank)       # before any pinned allocation
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def max_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.clips or {"cfg1": 2048, "cfg2": 1024, "cfg3": 256, "cfg4": 96}[args.workload]
    eng = TIMEngine(cfg, local_rank, args.dtype)
    sd = synth_state_dict(cfg, 0, "trained")
    eng.load_state_dict(sd)

    # ---- parity gate on a small seeded batch against the oracle (same weights) ----
    parity = None
    if rank == 0:
        from oracle.tim_oracle import TIMOracle
        pin = synth_inputs(cfg, 2, Qv, Qa, 4321, shared_queries=cfg.variant == "detection")
        ref = TIMOracle(cfg, sd, np.float32).forward(pin.get("vis"), pin.get("aud"), pin["times"], Qv, Qa, clip_chunk=1)
        te = eng.time_mlp(torch.from_numpy(pin["times"]).to(dev))
        o = eng.encoder(torch.from_numpy(pin["vis"]).to(dev) if "vis" in pin else None,
                        torch.from_numpy(pin["aud"]).to(dev) if "aud" in pin else None, te, Qv, Qa)
        errs = {k: rel_l2(o[k].cpu().numpy(), v) for k, v in ref.items() if v is not None and o.get(k) is not None}
        tol = {"fp32": 1e-5, "fp16": 1e-3, "bf16": 1e-2}[args.dtype]
        parity = {"max_rel_l2_vs_oracle": max(errs.values()), "tol": tol, "ok": max(errs.values()) <= tol}
        if not parity["ok"]:
            raise SystemExit(f"bench.py: parity gate failed: {errs}")

    # ---- synthetic inputs of the workload's shape, resident in HBM (and pinned host copies for e2e) ----
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    F = cfg.num_feats
    vis = torch.randn((B, F, cfg.visual_input_dim), generator=g, device=dev) if cfg.has_visual_input else None
    aud = torch.randn((B, F, cfg.audio_input_dim), generator=g, device=dev) if cfg.has_audio_input else None
    times = torch.from_numpy(synth_inputs(cfg, 1, Qv, Qa, 1234, shared_queries=cfg.variant == "detection")["times"]).to(dev)
    times = times.repeat(B, 1, 1).contiguous()
    times[:, cfg.F_tot:, 0] += 0.05 * torch.rand((B, times.shape[1] - cfg.F_tot), generator=g, device=dev)
    times[:, cfg.F_tot:, 1] += 0.08
    in_bytes = sum(t.numel() * 4 for t in (vis, aud, times) if t is not None)

    def step():
        te = eng.time_mlp(times)
        return eng.encoder(vis, aud, te, Qv, Qa)

    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = max_ranks(e0.elapsed_time(e1) / args.steps)
    launches = eng.launch_count - l0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = world * B * (Qv + Qa) / (ms * 1e-3)

    # ---- live per-kernel-class timing (CUDA events on the launching stream), same K steps ----
    eng.profile_begin()
    for _ in range(args.steps):
        step()
    prof = eng.profile_end()
    pk = peaks()
    gemm = prof["gemm"]
    gemm_tflops = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
    kinds = ("gemm_in_proj_linear1", "gemm_out_proj_linear2", "gemm_other")
    total_prof_ms = sum(c["ms"] for k, c in prof.items() if k != "gemm")
    by_kind = {k: {"tflops": prof[k]["flops"] / (prof[k]["ms"] * 1e-3) / 1e12 if prof[k]["ms"] > 0 else 0.0,
                   "frac": prof[k]["flops"] / (prof[k]["ms"] * 1e-3) / 1e12 / pk["tflops"] if prof[k]["ms"] > 0 else 0.0,
                   "ms_per_step": prof[k]["ms"] / args.steps, "launches_per_step": prof[k]["launches"] // args.steps} for k in kinds}
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("workload") == args.workload and tj.get("clips_per_step") == B:      # only for the configuration it was captured on
            traffic, traffic_note = tj["dram_bytes_per_launch"], f"{tj['what']}; {tj['source']}; algorithmic {tj['algorithmic_bytes_per_launch']:.3e} B"
    roofline = {"bound": "tensor", "kernel": "linear_umma2_kernel / linear_umma_kernel (tcgen05 GEMMs: every dense contraction of the step)",
                "achieved": gemm_tflops, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": gemm_tflops / pk["tflops"],
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": pk["source"],
                "launches_per_step": gemm["launches"] // args.steps, "avg_launch_ms": gemm["ms"] / max(1, gemm["launches"]),
                "share_of_step": gemm["ms"] / total_prof_ms if total_prof_ms else None,
                "class_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items() if k not in kinds},
                "by_gemm_kind": by_kind,
                "path_tflops": cfg.flops_fwd_per_clip(Qv, Qa) * B / (ms * 1e-3) / 1e12,
                "path_frac": cfg.flops_fwd_per_clip(Qv, Qa) * B / (ms * 1e-3) / 1e12 / pk["tflops"]}

    # ---- end to end: pinned host inputs -> plugin call -> pinned host outputs ----
    hv = vis.cpu().pin_memory() if vis is not None else None
    ha = aud.cpu().pin_memory() if aud is not None else None
    ht = times.cpu().pin_memory()
    # target chunk of the H2D / compute / D2H pipeline; the library tapers the ends and aligns chunk sizes to whole waves of GEMM
    # tiles (tim_forward_host). B // 3 measured best on cfg2 (27.8 ms vs 27.9 at B // 5 and 28.9 at B // 4, whose remainder
    # chunk of 34 clips runs at a quarter wave)
    chunk = args.chunk or max(1, B // 3)
    # what the reference's eval loop brings back to the host: the logits / regression outputs (recognition/scripts/test.py:
    # 106-131 reads output[0] only); the feature rows output[1] feed the drloc loss in training and stay on the device
    hout = eng._alloc_outputs(B, Qv, Qa, pinned=True, want_feats=False)
    for _ in range(2):
        _, up, down = eng.forward_host(hv, ha, ht, Qv, Qa, clips_per_chunk=chunk, out=hout)
    e2e_steps = max(3, args.steps // 4)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.forward_host(hv, ha, ht, Qv, Qa, clips_per_chunk=chunk, out=hout)
    torch.cuda.synchronize()
    e2e_s = max_ranks((time.perf_counter() - t0) / e2e_steps)
    barrier()
    e2e = {"value": world * B * (Qv + Qa) / e2e_s, "unit": UNIT, "h2d_bytes_per_step": up, "d2h_bytes_per_step": down,
           "ms_per_step": e2e_s * 1e3, "clips_per_chunk": chunk, "host_numa_node_of_rank0": numa_node,
           "d2h": "logits / regression outputs (what the reference eval loop reads, test.py:106-131); feature rows stay on device",
           "timing": "host wall clock around the blocking plugin call (outputs are in pinned host memory when it returns), "
                     "device synchronised on both sides, max over ranks"}

    # ---- the same end-to-end metric with the feature bank resident in HBM: the reference caches every
    # feature of the dataset in host RAM and gathers each clip's window on the host; with the bank on the device only the row
    # indices and interval times of a step cross PCIe (H2D), the logits still come back (D2H). Reported NEXT TO `e2e`, not as it.
    e2e_bank = None
    if cfg.has_visual_input and cfg.has_audio_input:
        vbank, abank = vis.reshape(B * F, -1), aud.reshape(B * F, -1)
        perm = torch.stack([torch.randperm(B * F, generator=torch.Generator().manual_seed(7 + i)) for i in range(2)])
        hvr = perm[0].view(B, F).contiguous().pin_memory()
        har = perm[1].view(B, F).contiguous().pin_memory()
        s_comp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        hres = {k: v for k, v in hout[0].items() if v is not None}

        def bank_step():
            moved = 0
            for b0 in range(0, B, chunk):
                b1 = min(B, b0 + chunk)
                with torch.cuda.stream(s_comp):
                    vr = hvr[b0:b1].to(dev, non_blocking=True)
                    ar = har[b0:b1].to(dev, non_blocking=True)
                    tt = ht[b0:b1].to(dev, non_blocking=True)
                    o = eng.encoder_indexed(vbank, vr, abank, ar, eng.time_mlp(tt), Qv, Qa, want_feats=False)
                    done = torch.cuda.Event()
                    done.record(s_comp)
                moved += (vr.numel() + ar.numel()) * 8 + tt.numel() * 4
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    for k, hbuf in hres.items():
                        rows = hbuf.shape[0] // B
                        hbuf[b0 * rows:b1 * rows].copy_(o[k], non_blocking=True)
                        o[k].record_stream(s_out)
            s_comp.synchronize(); s_out.synchronize()
            return moved

        for _ in range(2):
            up_bank = bank_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            bank_step()
        torch.cuda.synchronize()
        bank_s = max_ranks((time.perf_counter() - t0) / e2e_steps)
        barrier()
        e2e_bank = {"value": world * B * (Qv + Qa) / bank_s, "unit": UNIT, "ms_per_step": bank_s * 1e3, "h2d_bytes_per_step": up_bank,
                    "d2h_bytes_per_step": down, "bank": f"fp32 [{B * F}, {cfg.visual_input_dim}] + [{B * F}, {cfg.audio_input_dim}] resident in HBM, "
                    "windows gathered on the device (tim_encoder_fwd_indexed); random row permutation per modality"}

    if rank == 0:
        cb = None
        if not args.no_cpu_baseline:
            c = cpu_port(cfg, Qv, Qa, CPU_SAMPLE_CLIPS[args.workload], 8, 1)
            cb = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
                "data": "synthetic",
                "config": {"workload": f"{args.workload}: {wl_desc}", "clips_per_gpu_per_step": B, "queries_per_clip": Qv + Qa,
                           "operands": f"{args.dtype} tcgen05 operands, fp32 accumulate/softmax/LayerNorm/residual" if args.dtype != "fp32" else "fp32 CUDA cores",
                           "parallelism": f"dp{world} over clips, no data-path collective",
                           "l2": f"per-step inputs {in_bytes / 1e6:.0f} MB + activations exceed the 126 MB L2 (no flush needed)",
                           "weights": "synthetic trained-like (tim_b200.synth)"},
                "clips_per_sec": world * B / (ms * 1e-3), "tokens_per_sec": world * B * cfg.seq_len(Qv, Qa) / (ms * 1e-3),
                "gpu_launches": int(launches), "e2e": e2e, "e2e_resident_bank": e2e_bank, "roofline": roofline, "cpu_baseline": cb, "clocks": clocks, "parity": parity}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
