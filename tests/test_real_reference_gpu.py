"""The drop-in against the REAL reference module on a B200: tim_b200.patch_model() applied to the unmodified
`TIM` class (recognition and detection), compared with the same module run un-patched by eager PyTorch in fp32 (TF32 off) on the
same GPU - forward outputs in eval mode, and every parameter gradient in train() mode (dropout 0) against torch.autograd.

The reference package travels to the GPU box as baseline/_ref/ (git-ignored, staged by tools/stage_reference.py; in the build
container /root/reference is used directly). One subprocess per variant: both variants use the package name time_interval_machine.
"""
import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.refload import reference_root   # noqa: E402

TOL = {"eval": {"fp32": 1e-5, "fp16": 1e-3}, "train": {"fp32": 2e-5, "fp16": 1.5e-3}}


@pytest.mark.parametrize("mode", ["eval", "train"])
@pytest.mark.parametrize("variant", ["recognition", "detection"])
def test_patch_model_on_real_reference(variant, mode, tmp_path):
    if reference_root(variant) is None:
        pytest.skip("reference package not present (run tools/stage_reference.py in the build container: baseline/_ref travels to the GPU box)")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    out = tmp_path / f"{variant}_{mode}.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_ref_worker.py"), variant, mode, str(out)], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    rep = json.load(open(out))
    keep = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(keep):
        json.dump(rep, open(os.path.join(keep, f"real_reference_{variant}_{mode}.json"), "w"), indent=1)
    from tests.test_gpu_train import grad_tol, relu_gated
    worst = {}
    for case, rec in rep["cases"].items():
        assert rec, case
        real_width = case.startswith("cfg")
        for key, v in rec.items():
            dt, name = key.split("/", 1)
            if dt == "ref_autocast_fp16":          # the reference's own mixed-precision error: a yardstick, not a check
                worst[(case, "ref_autocast_fp16" + ("_gated" if relu_gated(name) else ""))] = max(
                    worst.get((case, "ref_autocast_fp16" + ("_gated" if relu_gated(name) else "")), 0.0), v["rel_l2"])
                continue
            if mode == "train" and name != "loss":
                tol = grad_tol(name, dt, real_width, variant == "detection")
                # no worse than 1.5 x what the reference's own fp16 autocast does to this tensor, where that is the larger bound
                if dt == "fp16":
                    tol = max(tol, 1.5 * rec.get(f"ref_autocast_fp16/{name}", {"rel_l2": 0.0})["rel_l2"])
            else:
                tol = TOL[mode][dt]
                # narrow test models (d_model = 64) sum few terms per output: their 16-bit bound is looser, the real widths hold the stated one
                if dt == "fp16" and not real_width:
                    tol *= 3.0
            tag = dt + ("_gated" if mode == "train" and relu_gated(name) else "")
            worst[(case, tag)] = max(worst.get((case, tag), 0.0), v["rel_l2"])
            assert v["rel_l2"] <= tol, f"{variant}/{mode}/{case}/{key}: rel-L2 {v['rel_l2']:.3e} > {tol:.1e}"
    print({f"{c}[{d}]": f"{e:.2e}" for (c, d), e in worst.items()})
