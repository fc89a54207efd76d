#!/usr/bin/env python
"""One tiny training step (forward + backward through the C ABI) per compute dtype: what compute-sanitizer runs over.
    python tools/train_smoke.py [case] [dtypes...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_gpu_train import cotangents, engine_grads   # noqa: E402
from tests.test_oracle_golden import MANIFEST   # noqa: E402
from tim_b200.config import TIMConfig   # noqa: E402
from tim_b200.synth import synth_inputs, synth_state_dict   # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "recog_av_small"
dts = sys.argv[2:] or ["fp16", "fp32"]
case = MANIFEST[name]
cfg = TIMConfig(**case["cfg"])
sd = synth_state_dict(cfg, case["weight_seed"], case["style"])
inp = synth_inputs(cfg, case["B"], case["Qv"], case["Qa"], case["input_seed"], shared_queries=case["shared_queries"])
for dt in dts:
    _, g = engine_grads(cfg, sd, inp, case["Qv"], case["Qa"], dt, lambda shapes: cotangents(name, shapes))
    print(dt, "gradient tensors", len(g), "max |g|", max(float(np.abs(v).max()) for v in g.values()), "finite", all(np.isfinite(v).all() for v in g.values()))
