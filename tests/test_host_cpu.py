"""CPU-side checks (no GPU): the C-ABI library loads and exports every declared symbol, fails loudly without a
device, and the host logic (config mirror, sharding over ranks with gloo) behaves."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tim_b200 import build, _lib
    build.build()                       # nvcc cross-compiles sm_100a without a GPU
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from tim_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tim_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|size_t|uint64_t|const char\*)\s+(tim_\w+)\s*\(", header, re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tim_abi_version() == _lib.ABI_VERSION


def test_config_struct_matches_header():
    from tim_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tim_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} tim_config;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n.strip() for decl in re.findall(r"int32_t\s+([^;]+);", body) for n in decl.split(",")]
    assert names == [f[0] for f in _lib.tim_config._fields_]
    body = re.search(r"typedef struct \{(.*?)\} tim_outputs;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    assert re.findall(r"float\*\s+(\w+);", body) == [f[0] for f in _lib.tim_outputs._fields_]


def test_seq_len_host_arithmetic(lib):
    from tim_b200.config import named_config
    from tim_b200.plugin import _c_config
    for name, S in (("cfg1", 70), ("cfg2", 200), ("cfg3", 528), ("cfg4", 2148)):
        cfg, Qv, Qa = named_config(name)
        cc = _c_config(cfg, "fp16")
        assert lib.tim_seq_len(C.byref(cc), Qv, Qa) == S == cfg.seq_len(Qv, Qa)


def test_no_device_is_a_loud_error(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tim_b200._lib import TimError
    from tim_b200.config import named_config
    from tim_b200.plugin import TIMEngine
    with pytest.raises(TimError, match="no CPU fallback"):
        TIMEngine(named_config("cfg1")[0], 0, "fp16")


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under tim_b200/ may import or execute it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle|oracle[./]tim_oracle|import_module\(.*oracle", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tim_b200")):
        for f in files:
            if f.endswith(".py"):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_shard_ranges():
    from tim_b200.dist import shard_range
    for n, w in ((10, 3), (8, 8), (5, 8), (1000, 7)):
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


def test_gloo_world2_sharded_forward_matches_single():
    """N>1 host path on CPU: two gloo ranks each take their clip shard (no data-path collective), results are
    gathered and must equal the unsharded run. The per-shard forward is the oracle here (no GPU in this test)."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tests", "_gloo_worker.py")], capture_output=True, text=True, timeout=300,
                       cwd=ROOT, env={**os.environ, "OMP_NUM_THREADS": "2"})
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_SHARD_OK" in r.stdout
    assert "GLOO_FLATGRAD_OK" in r.stdout              # flat gradient buffer + one all-reduce (training-leg groundwork)


@pytest.mark.parametrize("variant", ["recognition", "detection"])
def test_config_from_real_reference_model(variant):
    """Where the reference checkout exists (the build container, not the GPU box): instantiate the UNMODIFIED reference TIM on
    CPU and check that the drop-in reads its constructor arguments back correctly (plugin.config_from_model) and that every
    hot-path key exists in its state_dict with the expected shape. Own subprocess: recognition and detection share a package
    name."""
    import subprocess
    import sys
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference checkout not present on this machine")
    code = f"""
import sys
sys.path.insert(0, {ROOT!r})
from tools.make_golden import build_reference
from tim_b200.config import TIMConfig, hot_path_keys, state_dict_spec
from tim_b200.plugin import config_from_model
kw = dict(num_class=[[5, 7, 11], 3], visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4, num_layers=2, num_feats=6)
if {variant!r} == "detection":
    kw.update(num_class=[9, 4], include_verb_noun=False, data_modality="visual", variant="detection")
cfg = TIMConfig(**kw)
model = build_reference(cfg)
got = config_from_model(model)
assert got == cfg, (got, cfg)
sd, spec = model.state_dict(), state_dict_spec(cfg)
for k in hot_path_keys(cfg):
    assert tuple(sd[k].shape) == tuple(spec[k]), k
print("OK")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_forward_host_chunk_schedule(lib):
    """The chunk sizes tim_forward_host cuts a batch into (pure host arithmetic of the library, no device): they add up to B, none
    is empty or larger than the target, and on the 16-bit path with E a multiple of 256 every chunk except the remainder is a whole
    number of GEMM waves (row tiles a multiple of num_sms / gcd(num_sms, E / 256), rounded down to whole clips); the ends taper."""
    import ctypes as C
    import math
    import random

    def sched(B, cpc, rows, E, sms, sixteen):
        buf = (C.c_int * 4096)()
        n = lib.tim_host_chunk_schedule(B, cpc, rows, E, sms, int(sixteen), buf, 4096)
        assert 0 < n <= 4096, (B, cpc, rows, E, sms, sixteen, n)
        return list(buf[:n])

    # the bench's end-to-end call: cfg2, 1024 clips, target B // 3
    assert sched(1024, 341, 200, 1024, 148, True) == [47, 142, 315, 331, 142, 47]
    assert sched(1024, 256, 200, 1024, 148, True) == [47, 94, 34, 236, 236, 236, 94, 47]
    # fp32 mode / widths the pair kernel does not tile: plain tapered chunks
    assert sched(96, 24, 2148, 1024, 148, False) == [6, 12, 12, 24, 24, 12, 6]
    assert sched(4, 0, 200, 1024, 148, True) == [4]
    assert sched(5, 99, 13, 128, 148, True) == [5]
    rnd = random.Random(0)
    for _ in range(3000):
        B = rnd.randint(1, 40000)
        cpc = rnd.choice([0, rnd.randint(1, B), rnd.randint(1, 2 * B), B // 3 + 1, B // 4 + 1])
        rows = rnd.choice([13, 24, 100, 200, 528, 2148])
        E = rnd.choice([128, 256, 512, 1024, 1536, 2048, 96])
        sms = rnd.choice([148, 132, 8])
        sixteen = rnd.random() < 0.8
        ch = sched(B, cpc, rows, E, sms, sixteen)
        target = B if (cpc <= 0 or cpc > B) else cpc
        assert sum(ch) == B and min(ch) > 0 and max(ch) <= target, (B, cpc, rows, E, sms, sixteen, ch)
        if sixteen and E % 256 == 0:
            unit_tiles = sms // math.gcd(sms, E // 256)
            unit = unit_tiles * 256 // rows
            if unit >= 8 and target >= unit and ((target + 1) * rows - 1) // (unit_tiles * 256) >= 1:
                k = ((target + 1) * rows - 1) // (unit_tiles * 256)      # the largest k whose k units of waves fit the target
                body = k * unit_tiles * 256 // rows          # a full-size chunk: k units of waves, rounded down to whole clips
                assert max(ch) <= body and (B < body or max(ch) == body)
                assert -(-body * rows // 256) <= k * unit_tiles    # its row tiles fit k units exactly
                aligned = [c for c in ch if any(c == j * unit_tiles * 256 // rows for j in range(1, k + 1))]
                assert len(aligned) >= len(ch) - 1          # at most one remainder chunk


def test_postprocess_without_a_device_is_a_loud_error(lib):
    """tim_b200.postprocess has no CPU path either: without a CUDA device every entry point raises instead of computing on the
    host (and an explicit CPU device is refused)."""
    import numpy as np
    import torch
    from tim_b200 import postprocess as pp
    segs, scores, cls = np.array([[0.0, 1.0], [0.5, 1.5]], np.float32), np.array([0.9, 0.8], np.float32), np.zeros(2, np.int64)
    with pytest.raises(RuntimeError, match="CUDA device only"):
        pp.batched_nms(segs, scores, cls, 0.1, 0.001, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA device only"):
        pp.threshold_detections(torch.zeros((2, 3)), torch.zeros((2, 2), dtype=torch.float64), 0.03, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA device only"):
        pp.decode_predictions(torch.zeros((2, 3)), torch.zeros((2, 2)), torch.zeros(1, dtype=torch.float64), 30.0, 1.0, device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            pp.batched_nms(segs, scores, cls, 0.1, 0.001)
    with pytest.raises(NotImplementedError):
        pp.batched_nms(segs, scores, cls, 0.1, 0.001, multi_class=False)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (CPU only, no GPU needed): ONE JSON line on stdout carrying the keys the driver reads - the same
    metric / unit / config as the GPU arm, `impl: reference`, a `cpu_baseline` describing the run and an `e2e` object without
    device copies; under torchrun only rank 0 prints (checked here by running a second process as RANK=1)."""
    import json
    env = {**os.environ, "OMP_NUM_THREADS": "1"}           # what torchrun hands its workers
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "clips_x_queries_per_sec" and d["unit"] == "clips*queries/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0)), "the CPU arm must use every host thread even under OMP_NUM_THREADS=1"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1",
                         "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                        env={**env, "RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""
