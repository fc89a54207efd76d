#!/usr/bin/env python
"""1-vs-N-GPU gradient equality of the training leg (torchrun, one rank per GPU):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_grad_check.py [--out gpurun_out/ddp_grad_check.json]

Every rank runs the autograd drop-in (patch_model on a module with the reference's parameter tree, train() mode) on ITS shard of one
seeded batch with loss = sum <output, cotangent>; the backward queues the library's single all-reduce (ncclAllReduce, average) over
the flat gradient buffer exactly as in training. N x the averaged gradient must equal the gradient of the whole batch computed on one
GPU through the engine API without any collective. Also times the all-reduce of the real cfg2 / cfg4 gradient buffers (233 / 221 MB).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests._fake_tim import FakeTIM   # noqa: E402
from tests.test_gpu_train import cotangents, engine_grads   # noqa: E402
from tim_b200.config import named_config   # noqa: E402
from tim_b200.dist import shard_range   # noqa: E402
from tim_b200.plugin import TIMEngine, patch_model   # noqa: E402
from tim_b200.synth import rel_l2, synth_inputs, synth_state_dict   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report = {"world": world, "cases": {}}
    for dt, tol in (("fp32", 1e-5), ("fp16", 2e-3)):
        cfg, Qv, Qa = named_config("cfg1")
        B = 4 * world
        sd = synth_state_dict(cfg, 0, "trained")
        inp = synth_inputs(cfg, B, Qv, Qa, 321)
        lo, hi = shard_range(B, rank, world)
        model = patch_model(FakeTIM(cfg, sd).to(dev).train(), compute_dtype=dt)
        vis, aud, times = (torch.from_numpy(inp[k][lo:hi]).to(dev) for k in ("vis", "aud", "times"))
        te = model(times, "time_mlp")
        (verb, noun, action, audio), feats = model([vis, aud], "encoder", te, Qv, Qa)
        outs = dict(verb=verb, noun=noun, action=action, audio=audio, feats=feats)
        full_shapes = {k: (v.shape[0] // (hi - lo) * B,) + tuple(v.shape[1:]) for k, v in outs.items()}
        cot = cotangents("ddp", full_shapes)
        loss = 0.0
        for k, v in outs.items():
            rows = v.shape[0] // (hi - lo)
            c = torch.from_numpy(cot[k][lo * rows:hi * rows].astype(np.float32)).to(dev)
            loss = loss + (v * c).sum()
        loss.backward()                                   # ... and the queued callback all-reduces the flat buffer
        torch.cuda.synchronize()
        flat = model._tim_b200.flat
        if rank == 0:
            _, gref = engine_grads(cfg, sd, inp, Qv, Qa, dt, lambda shapes: cot)
            worst = ("", 0.0)
            for k, g in gref.items():
                got = flat.view(k).view(g.shape).cpu().numpy() * world
                e = rel_l2(got, g)
                if e > worst[1]:
                    worst = (k, e)
            report["cases"][dt] = {"worst_key": worst[0], "worst_rel_l2": worst[1], "tol": tol, "ok": worst[1] <= tol,
                                   "flat_buffer_mb": flat.buffer.numel() * 4 / 1e6}
        dist.barrier(device_ids=[local])
        del model
    # the collective alone on the real gradient-buffer sizes
    for name in ("cfg2", "cfg4"):
        cfg, _, _ = named_config(name)
        eng = TIMEngine(cfg, local, "fp16")
        eng.comm_init()
        n = sum(int(np.prod(s)) for s in eng._spec.values())
        buf = torch.ones(n, dtype=torch.float32, device=dev) * (rank + 1)
        for _ in range(3):
            eng.allreduce(buf)
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.allreduce(buf)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            report["cases"][f"allreduce_{name}"] = {"bytes": n * 4, "ms": float(t.item()), "algbw_gbs": n * 4 / (float(t.item()) * 1e-3) / 1e9,
                                                     "busbw_gbs": n * 4 / (float(t.item()) * 1e-3) / 1e9 * 2 * (world - 1) / world}
        eng.close()
    if rank == 0:
        print(json.dumps(report))
        if args.out:
            json.dump(report, open(args.out, "w"), indent=1)
        assert all(c.get("ok", True) for c in report["cases"].values()), report
    import threading
    w = threading.Timer(45.0, lambda: os._exit(0))       # results are out; never hang in the process-group teardown
    w.daemon = True
    w.start()
    try:
        dist.destroy_process_group()
    except Exception:      # peers may already be gone (their watchdog): the results are out
        pass
    w.cancel()


if __name__ == "__main__":
    main()
