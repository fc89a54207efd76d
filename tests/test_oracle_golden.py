"""Pins oracle/tim_oracle.py to the outputs of the real reference (tests/golden/*.npz, minted by
tools/make_golden.py from /root/reference on CPU fp32). Runs without a GPU and without the reference."""
import json
import os

import numpy as np
import pytest

from oracle.tim_oracle import TIMOracle
from tim_b200.config import TIMConfig, state_dict_spec, named_config
from tim_b200.synth import synth_inputs, synth_state_dict, rel_l2

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]
OUT_KEYS = ("verb", "noun", "action", "audio", "reg_v", "reg_a")


def load_case(name):
    c = MANIFEST[name]
    cfg = TIMConfig(**c["cfg"])
    sd = synth_state_dict(cfg, c["weight_seed"], c["style"])
    inp = synth_inputs(cfg, c["B"], c["Qv"], c["Qa"], c["input_seed"], shared_queries=c["shared_queries"])
    gold = dict(np.load(os.path.join(GOLD, name + ".npz")))
    if c["pyramid"]:
        inp["times"][:, cfg.F_tot:] = pyramid_queries()
    return cfg, sd, inp, gold, c


def pyramid_queries(query_size=0.01):
    """detection/.../tim.py:144-155 generate_queries, restated in numpy (float32 like torch.arange)."""
    out = []
    while query_size < 1.0:
        # torch.arange(0., 1., step) has ceil((1-0)/step) elements
        n = int(np.ceil(1.0 / (query_size / 2)))
        st = (np.arange(n, dtype=np.float64) * (query_size / 2)).astype(np.float32)
        en = st + np.float32(query_size)
        out.append(np.round(np.stack([st, en], -1), 3))
        query_size *= 2
    return np.concatenate(out, 0)[None].astype(np.float32)


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_oracle_matches_reference_golden(name):
    cfg, sd, inp, gold, c = load_case(name)
    o = TIMOracle(cfg, sd, np.float32).forward(inp.get("vis"), inp.get("aud"), inp["times"], c["Qv"], c["Qa"])
    for k in OUT_KEYS:
        assert (o[k] is None) == (k not in gold), k
    for k, g in gold.items():
        assert o[k].shape == g.shape, (k, o[k].shape, g.shape)
        assert o[k].dtype == np.float32
        e = rel_l2(o[k], g)
        assert e <= 1e-5, f"{name}/{k}: rel-L2 {e:.3e} > 1e-5 (fp32 tolerance of BASELINE.json)"


def test_pyramid_has_399_queries():
    q = pyramid_queries()
    assert q.shape == (1, 399, 2)          # SURVEY.md §8(a9) [probe]


def test_mask_structure():
    cfg, Qv, Qa = named_config("cfg1")
    o = TIMOracle(cfg, {}, np.float32)
    S = cfg.seq_len(Qv, Qa)
    m = o.mask(S)
    assert S == 70 and m.shape == (S, S)
    assert not m[:, :cfg.F_tot].any()                       # every token sees all feature tokens
    assert not m[np.arange(S), np.arange(S)].any()          # ... and itself
    off = m[cfg.F_tot:, cfg.F_tot:]
    assert (off | np.eye(S - cfg.F_tot, dtype=bool)).all()  # queries never see other queries


def test_query_subset_invariance():
    """a query's logits do not depend on which other queries are present."""
    cfg, sd, inp, gold, c = load_case("recog_av_small")
    o = TIMOracle(cfg, sd, np.float64)
    full = o.forward(inp["vis"], inp["aud"], inp["times"], c["Qv"], c["Qa"])
    F = cfg.F_tot
    keep = np.r_[np.arange(F), F, F + c["Qv"]]              # first visual + first audio query only
    sub = o.forward(inp["vis"], inp["aud"], inp["times"][:, keep], 1, 1)
    B = c["B"]
    np.testing.assert_allclose(sub["action"], full["action"].reshape(B, c["Qv"], -1)[:, 0], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(sub["audio"], full["audio"].reshape(B, c["Qa"], -1)[:, 0], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(sub["feats"], full["feats"], rtol=1e-9, atol=1e-11)


def test_state_dict_spec_counts():
    cfg, _, _ = named_config("cfg1")
    n = sum(int(np.prod(s)) for s in state_dict_spec(cfg).values())
    assert n == 16_304_280                                  # SURVEY.md §8(c) [probe]
    cfg4, _, _ = named_config("cfg4")
    assert len(state_dict_spec(cfg4)) == 105                # 105 keys for 6L visual det
    assert sum(int(np.prod(s)) for s in state_dict_spec(cfg4).values()) > 50_000_000


def test_flop_formula_matches_survey():
    cfg2, Qv, Qa = named_config("cfg2")
    assert abs(cfg2.flops_fwd_per_clip(Qv, Qa) / 1e9 - 21.17) < 0.05
    cfg3, Qv, Qa = named_config("cfg3")
    assert abs(cfg3.flops_fwd_per_clip(Qv, Qa) / 1e9 - 123.7) < 0.5
    cfg4, Qv, Qa = named_config("cfg4")
    assert abs(cfg4.flops_fwd_per_clip(Qv, Qa) / 1e9 - 227.7) < 0.5


# ---------------------------------------------------------------- detection query labelling
LABEL_CASES = {"vn": [5, 7, 11], "act": [None, None, 9], "aud": [4]}


@pytest.mark.parametrize("name", sorted(LABEL_CASES))
def test_label_oracle_matches_reference_golden(name):
    """oracle/label_oracle.py against the outputs of the reference's label_queries / assign_positive_labels
    (tests/golden/label_queries.npz, minted by tools/make_golden_labels.py): bit-exact, NaN-aware."""
    from oracle.label_oracle import label_queries, smooth_labels
    g = np.load(os.path.join(GOLD, "label_queries.npz"))
    thr, sm = g[f"{name}_meta"]
    t, ids, iou = label_queries(g[f"{name}_queries"], g[f"{name}_gt"], g[f"{name}_labels"], thr)
    assert np.array_equal(t, g[f"{name}_targets"])
    assert np.array_equal(iou, g[f"{name}_ious"], equal_nan=True)
    for k, C_ in enumerate(LABEL_CASES[name]):
        ref = g[f"{name}_smooth{k}"]
        if C_ is None:
            assert ref.size == 0
            continue
        assert np.array_equal(smooth_labels(ids[:, k if name != "aud" else 0], C_, sm), ref)


# ---------------------------------------------------------------------------------------------------------------------
# detection post-processing: 1-D (soft-)NMS
# ---------------------------------------------------------------------------------------------------------------------
def nms_golden():
    return np.load(os.path.join(GOLD, "nms.npz"))


def nms_case(g, name):
    """(kind, segs, scores, cls | None, params) of one golden case."""
    prm = {k.split("/p_")[1]: g[k].item() for k in g.files if k.startswith(f"{name}/p_")}
    cls = g[f"{name}/cls"] if f"{name}/cls" in g.files else None
    return str(g[f"{name}/kind"]), g[f"{name}/segs"], g[f"{name}/scores"], cls, prm


NMS_CASES = [str(n) for n in np.load(os.path.join(GOLD, "nms.npz"))["names"]]
NMS_SCORE_RTOL = 1e-6          # glibc expf (reference) vs correctly rounded exp (oracle, device): at most an ulp or two


@pytest.mark.parametrize("name", NMS_CASES)
def test_nms_oracle_matches_reference_golden(name):
    """oracle/nms_oracle.py against the outputs of the reference's own compiled extension and its batched_nms driver
    (tests/golden/nms.npz, minted by tools/make_golden_nms.py): pick order and indices exact, segments exact, decayed scores to
    NMS_SCORE_RTOL. For hard NMS with equal scores the reference's pick among them follows torch's unstable sort, so that case
    pins the kept score sequence and the kept count instead of the indices."""
    from oracle import nms_oracle as o
    g = nms_golden()
    kind, segs, scores, cls, prm = nms_case(g, name)
    if kind == "soft":
        inds, dets = o.softnms_1d(segs, scores, **prm)
        assert np.array_equal(inds, g[f"{name}/inds"])
        ref = g[f"{name}/dets"]
        assert np.array_equal(dets[:, :2], ref[:, :2])
        np.testing.assert_allclose(dets[:, 2], ref[:, 2], rtol=NMS_SCORE_RTOL, atol=0)
    elif kind == "nms":
        inds = o.nms_1d(segs, scores, **prm)
        if name.endswith("ties"):
            assert np.array_equal(scores[inds], scores[g[f"{name}/inds"]])
        else:
            assert np.array_equal(inds, g[f"{name}/inds"])
    else:
        s, p, c = o.batched_nms(segs, scores, cls, **prm)
        assert np.array_equal(s, g[f"{name}/out_segs"]) and np.array_equal(c, g[f"{name}/out_cls"])
        np.testing.assert_allclose(p, g[f"{name}/out_scores"], rtol=NMS_SCORE_RTOL, atol=0)


def test_nms_oracle_matches_compiled_reference_random():
    """More seeded cases against oracle/_ref/nms_1d_cpu.so (the reference's source compiled by oracle/build_ref.py) when it is
    present; the golden file above is the pin that travels."""
    from oracle import build_ref, nms_oracle as o
    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref/nms_1d_cpu.so not built (reference tree absent)")
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(7)
    for trial in range(12):
        n = int(rng.integers(1, 400))
        c = rng.uniform(0, 30, 5)
        start = c[rng.integers(0, 5, n)] + rng.normal(0, 0.5, n)
        segs = np.stack([start, start + np.abs(rng.normal(2, 0.5, n)) + 0.05], 1).astype(np.float32)
        scores = rng.uniform(0.01, 1, n).astype(np.float32)
        method, mins, thr = trial % 3, float(rng.choice([0.001, 0.05])), float(rng.choice([0.1, 0.5]))
        dets = torch.zeros((n, 3))
        ref_inds = mod.softnms(torch.from_numpy(segs), torch.from_numpy(scores), dets, thr, 0.25, mins, method).numpy()
        inds, d = o.softnms_1d(segs, scores, thr, 0.25, mins, method)
        assert np.array_equal(inds, ref_inds), trial
        np.testing.assert_allclose(d, dets[:len(inds)].numpy(), rtol=NMS_SCORE_RTOL, atol=0)
        assert np.array_equal(o.nms_1d(segs, scores, thr), mod.nms(torch.from_numpy(segs), torch.from_numpy(scores), thr).numpy())


def test_parallel_deletion_equals_sequential_scan():
    """The identity tim_b200/csrc/nms.cu relies on: the reference deletes inside its decay scan by moving the LAST live entry into
    the hole and re-examining it (nms_cpu.cpp:147-156); with n' survivors, that leaves the k-th hole below n' (ascending) holding
    the k-th surviving entry at or above n' counted from the end. Literal transcription vs the closed form, random flags."""
    rng = np.random.default_rng(3)
    for _ in range(3000):
        n = int(rng.integers(1, 48))
        i = int(rng.integers(0, n))
        vals = list(rng.permutation(n))
        pr = float(rng.choice([0.05, 0.3, 0.7, 1.0]))
        removed = {v for q, v in enumerate(vals) if q > i and rng.random() < pr}
        a, m, pos = list(vals), n, i + 1
        while pos < m:                                     # the reference's scan
            if a[pos] in removed:
                a[pos] = a[m - 1]
                m -= 1
                pos -= 1
            pos += 1
        b = list(vals)
        n2 = n - len(removed)
        holes = [q for q in range(i + 1, n2) if b[q] in removed]
        movers = [q for q in range(n2, n) if b[q] not in removed]
        assert len(holes) == len(movers)
        for k, h in enumerate(holes):
            b[h] = vals[movers[len(movers) - 1 - k]]
        assert a[:m] == b[:n2]


def detpost_golden():
    return np.load(os.path.join(GOLD, "detpost.npz"))


def test_detpost_oracle_matches_reference_golden():
    """oracle/nms_oracle.py's decode / threshold / format restatements against the UNMODIFIED reference chain
    (tests/golden/detpost.npz, tools/make_golden_detpost.py: FeatureMeter.update + finalize_metrics, then main() of
    eval_detection/format_predictions.py with the reference's nms.py and compiled extension): proposals bit-exact (float64),
    sigmoid scores to 1e-6 (exp rounding), final detections per video: classes and 3-decimal segments exact, scores to 1e-6."""
    from oracle import nms_oracle as o
    g = detpost_golden()
    window_size, thr, sigma, _ = g["params"]
    assert g["v_proposals"].dtype == np.float64 and g["action"].dtype == np.float32
    preds, props = [], []
    for b in range(2):
        p, pr = o.decode_predictions(g[f"b{b}/logits"], g[f"b{b}/reg"], g[f"b{b}/window_start"], window_size,
                                     g[f"b{b}/queries"].max())
        preds.append(p)
        props.append(pr)
    assert np.array_equal(np.concatenate(props), g["v_proposals"])
    np.testing.assert_allclose(np.concatenate(preds), g["action"], rtol=NMS_SCORE_RTOL, atol=0)
    res = o.format_predictions(g["action"], g["v_proposals"], g["video_ids"], thr, sigma)
    assert sorted(res) == sorted(str(v) for v in g["result_videos"])
    for v in g["result_videos"]:
        s, p, c = res[str(v)]
        assert np.array_equal(c, g[f"result/{v}/action"])
        assert np.array_equal(np.array([[round(float(a), 3), round(float(b), 3)] for a, b in s]), g[f"result/{v}/segment"])
        np.testing.assert_allclose(p.astype(np.float64), g[f"result/{v}/score"], rtol=NMS_SCORE_RTOL, atol=0)


# ---------------------------------------------------------------------------------------------------------------------
# backward oracle (groundwork for the training leg)
# ---------------------------------------------------------------------------------------------------------------------
GRAD_CASES = [str(n) for n in np.load(os.path.join(GOLD, "grads.npz"))["cases"]]


@pytest.mark.parametrize("name", GRAD_CASES)
def test_backward_oracle_matches_reference_autograd(name):
    """oracle/tim_oracle_bwd.py (numpy reverse mode, float64) against torch.autograd over the UNMODIFIED reference
    (tests/golden/grads.npz, tools/make_golden_grads.py): for every parameter that receives a gradient and for both feature
    inputs, the L2 norm, the sum and 512 sampled entries of d L / d tensor agree to 1e-9 relative (of the tensor's largest sampled
    magnitude); parameters autograd leaves without gradient get none here either."""
    import zlib
    from oracle.tim_oracle_bwd import TIMOracleGrad
    g = np.load(os.path.join(GOLD, "grads.npz"))
    case = MANIFEST[name]
    cfg = TIMConfig(**case["cfg"])
    sd = synth_state_dict(cfg, case["weight_seed"], case["style"])
    inp = synth_inputs(cfg, case["B"], case["Qv"], case["Qa"], case["input_seed"], shared_queries=case["shared_queries"])
    cot = {}
    for key in g.files:
        if key.startswith(f"{name}/out_shape/"):
            k = key.rsplit("/", 1)[1]
            cot[k] = np.random.default_rng(zlib.crc32(f"{name}/{k}".encode())).standard_normal(tuple(g[key]))
    _, grads = TIMOracleGrad(cfg, sd, np.float64).forward_backward(inp.get("vis"), inp.get("aud"), inp["times"], case["Qv"], case["Qa"], cot)
    want = [str(k) for k in g[f"{name}/keys"]]
    assert sorted(grads) == want, (sorted(set(grads) ^ set(want)))
    for k in want:
        flat = np.asarray(grads[k], np.float64).reshape(-1)
        idx = np.sort(np.random.default_rng(zlib.crc32(f"idx/{name}/{k}".encode())).choice(flat.size, size=min(512, flat.size), replace=False))
        norm, total = g[f"{name}/stat/{k}"]
        vals = g[f"{name}/vals/{k}"]
        scale = max(np.abs(vals).max(), 1e-30)
        assert np.abs(flat[idx] - vals).max() <= 1e-9 * scale, k
        assert abs(np.sqrt((flat * flat).sum()) - norm) <= 1e-9 * max(norm, 1e-30), k
        assert abs(flat.sum() - total) <= 1e-9 * max(norm * np.sqrt(flat.size), 1e-30), k


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_torch_op_port_matches_reference_golden(name):
    """oracle/tim_oracle_torch.py - the TIMED CPU arm of bench.py, which restates the reference with the PyTorch CPU calls the
    reference's modules make - against the golden vectors of the real reference: same tolerance as the numpy oracle (it is in fact
    bit-identical to the reference on the machine that minted them), so what the bench times on the host computes the reference's
    outputs."""
    pytest.importorskip("torch")
    from oracle.tim_oracle_torch import TIMOracleTorch
    cfg, sd, inp, gold, case = load_case(name)
    Qv = inp["times"].shape[1] - cfg.F_tot if case["pyramid"] else case["Qv"]
    out = TIMOracleTorch(cfg, sd).forward(inp.get("vis"), inp.get("aud"), inp["times"], Qv, case["Qa"])
    for k, ref in gold.items():
        assert out[k] is not None, k
        assert rel_l2(out[k], ref) <= 2e-6, (k, rel_l2(out[k], ref))


@pytest.mark.parametrize("B,Ft,Qt,H,hd", [(2, 12, 7, 2, 16), (3, 10, 0, 2, 8), (1, 5, 9, 1, 32)])
def test_two_stream_attention_oracles_match_dense_masked_attention(B, Ft, Qt, H, hd):
    """oracle/kernel_oracles.py (forward and backward of the attention core in the library's two-stream layout, exploiting the mask's
    structure) against the reference's formulation - a dense S x S masked softmax per clip (tim.py:161-166 mask) differentiated by
    torch.autograd in float64: outputs and d L / d qkv agree to 1e-12."""
    torch = pytest.importorskip("torch")
    from oracle.kernel_oracles import LN2, attention_bwd_two_stream, attention_fwd_two_stream
    rng = np.random.default_rng(Ft * 7 + Qt)
    E, S = H * hd, Ft + Qt
    M = B * S
    qkv = rng.standard_normal((M, 3 * E))
    dout = rng.standard_normal((M, E))
    out, _ = attention_fwd_two_stream(qkv, B, Ft, Qt, H, hd)
    dqkv = attention_bwd_two_stream(qkv, dout, B, Ft, Qt, H, hd)

    x = torch.from_numpy(qkv).requires_grad_(True)
    mask = torch.ones(S, S, dtype=torch.bool)
    mask[:, :Ft] = False
    mask.fill_diagonal_(False)
    ref = torch.zeros(M, E, dtype=torch.float64)
    for b in range(B):
        rows = torch.cat([torch.arange(b * Ft, (b + 1) * Ft), B * Ft + torch.arange(b * Qt, (b + 1) * Qt)])
        xb = x[rows]
        q, k, v = (xb[:, i * E:(i + 1) * E].reshape(S, H, hd).transpose(0, 1) for i in range(3))
        p = torch.softmax(((q @ k.transpose(1, 2)) * LN2).masked_fill(mask[None], float("-inf")), -1)
        ref = ref.index_add(0, rows, (p @ v).transpose(0, 1).reshape(S, E))
    (ref * torch.from_numpy(dout)).sum().backward()
    assert np.abs(out - ref.detach().numpy()).max() <= 1e-12 * max(1.0, np.abs(out).max())
    assert np.abs(dqkv - x.grad.numpy()).max() <= 1e-12 * max(1.0, np.abs(dqkv).max())


def test_layernorm_bwd_rows_matches_autograd():
    torch = pytest.importorskip("torch")
    from oracle.kernel_oracles import layernorm_bwd_rows
    rng = np.random.default_rng(2)
    x, dy, g, b = rng.standard_normal((37, 64)), rng.standard_normal((37, 64)), rng.standard_normal(64), rng.standard_normal(64)
    tx, tg, tb = (torch.from_numpy(a).requires_grad_(True) for a in (x, g, b))
    (torch.nn.functional.layer_norm(tx, (64,), tg, tb, 1e-5) * torch.from_numpy(dy)).sum().backward()
    dx, dg, db = layernorm_bwd_rows(dy, x, g)
    for got, want in ((dx, tx.grad), (dg, tg.grad), (db, tb.grad)):
        assert np.abs(got - want.numpy()).max() <= 1e-12 * max(1.0, np.abs(got).max())
