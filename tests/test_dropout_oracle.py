"""Dropout of the training leg, CPU side: the numpy restatement of the library's counter-based masks (oracle/tim_oracle_bwd.py:
drop_mask) and its placement at the reference's six nn.Dropout sites, pinned to torch.autograd over the UNMODIFIED reference run
with those masks substituted at its own dropout calls (tests/golden/grads_dropout.npz, tools/make_golden_grads_dropout.py)."""
import json
import os
import zlib

import numpy as np
import pytest

from tim_b200.config import TIMConfig
from tim_b200.synth import synth_inputs, synth_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")
_G = np.load(os.path.join(GOLD, "grads_dropout.npz"))
DROP_MANIFEST = json.loads(str(_G["manifest"]))
DROP_CASES = [str(n) for n in _G["cases"]]


def drop_case(name):
    """(cfg, state_dict, inputs, Qv, Qa, cotangents, dropout kwargs) of a tests/golden/grads_dropout.npz case."""
    m = DROP_MANIFEST
    case = m["cases"][name]
    cfg = TIMConfig(**case["cfg"])
    sd = synth_state_dict(cfg, m["weight_seed"], "trained")
    inp = synth_inputs(cfg, case["B"], case["Qv"], case["Qa"], m["input_seed"], shared_queries=cfg.variant == "detection")
    cot = {}
    for key in _G.files:
        if key.startswith(f"{name}/out_shape/"):
            k = key.rsplit("/", 1)[1]
            cot[k] = np.random.default_rng(zlib.crc32(f"{name}/{k}".encode())).standard_normal(tuple(_G[key]))
    drop = {"p_feat": m["p_feat"], "p_seq": m["p_seq"], "p_enc": m["p_enc"], "seed": m["seed"]}
    return cfg, sd, inp, case["Qv"], case["Qa"], cot, drop


def test_mask_statistics_and_determinism():
    from oracle.tim_oracle_bwd import drop_mask
    e = np.arange(1 << 20)
    for p in (0.1, 0.5):
        m = drop_mask(e, p, 1234, 3, 2)
        keep = (m > 0).mean()
        assert abs(keep - (1.0 - p)) < 3e-3, (p, keep)
        assert abs(m.mean() - 1.0) < 6e-3                      # kept values are scaled by 1 / (1 - p): the mask has mean 1
        assert np.array_equal(m, drop_mask(e, p, 1234, 3, 2))
        # different layer / site / seed -> a different, uncorrelated mask
        for other in (drop_mask(e, p, 1234, 3, 3), drop_mask(e, p, 1234, 4, 2), drop_mask(e, p, 1235, 3, 2)):
            agree = ((m > 0) == (other > 0)).mean()
            assert abs(agree - (p * p + (1 - p) * (1 - p))) < 5e-3, agree
        # neighbours (the two halves of one hash word) are uncorrelated too
        k = m > 0
        assert abs((k[0::2] & k[1::2]).mean() - (1 - p) ** 2) < 5e-3
    assert np.all(drop_mask(e[:100], 0.0, 1, 1) == 1.0)


@pytest.mark.parametrize("name", DROP_CASES)
def test_backward_oracle_with_dropout_matches_reference_autograd(name):
    """Outputs and every gradient of the oracle with dropout = the reference's autograd with the same masks at its own dropout
    modules (1e-9 relative, float64)."""
    from oracle.tim_oracle_bwd import TIMOracleGrad
    cfg, sd, inp, Qv, Qa, cot, drop = drop_case(name)
    out, grads = TIMOracleGrad(cfg, sd, np.float64).forward_backward(inp.get("vis"), inp.get("aud"), inp["times"], Qv, Qa, cot, dropout=drop)
    for k in cot:
        ref = _G[f"{name}/out/{k}"]
        assert np.abs(np.asarray(out[k]).reshape(ref.shape) - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-30), k
    want = [str(k) for k in _G[f"{name}/keys"]]
    assert sorted(grads) == want, sorted(set(grads) ^ set(want))
    for k in want:
        flat = np.asarray(grads[k], np.float64).reshape(-1)
        idx = np.sort(np.random.default_rng(zlib.crc32(f"idx/{name}/{k}".encode())).choice(flat.size, size=min(512, flat.size), replace=False))
        norm, total = _G[f"{name}/stat/{k}"]
        vals = _G[f"{name}/vals/{k}"]
        scale = max(np.abs(vals).max(), 1e-30)
        assert np.abs(flat[idx] - vals).max() <= 1e-9 * scale, k
        assert abs(np.sqrt((flat * flat).sum()) - norm) <= 1e-9 * max(norm, 1e-30), k
        assert abs(flat.sum() - total) <= 1e-9 * max(norm * np.sqrt(flat.size), 1e-30), k


def test_dropout_changes_the_result():
    """Guard against a vacuous pin: with the masks the outputs differ from the p = 0 forward by O(1)."""
    from oracle.tim_oracle_bwd import TIMOracleGrad
    cfg, sd, inp, Qv, Qa, cot, drop = drop_case("recog_av_small")
    o = TIMOracleGrad(cfg, sd, np.float64)
    a, _ = o.forward_backward(inp.get("vis"), inp.get("aud"), inp["times"], Qv, Qa, cot, dropout=drop)
    b, _ = o.forward_backward(inp.get("vis"), inp.get("aud"), inp["times"], Qv, Qa, cot)
    assert np.linalg.norm(a["action"] - b["action"]) > 0.1 * np.linalg.norm(b["action"])
