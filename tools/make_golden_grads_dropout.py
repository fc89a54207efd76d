#!/usr/bin/env python
"""Mints tests/golden/grads_dropout.npz: torch.autograd gradients of the UNMODIFIED reference forward WITH dropout active at all
six of its nn.Dropout sites (float64, CPU) - the pin of the dropout placement in oracle/tim_oracle_bwd.py and, through it, of
tim_set_dropout in the CUDA training leg.

    python tools/make_golden_grads_dropout.py          (build container only; one subprocess per variant)

torch's own dropout masks come from its Philox stream and cannot be reproduced by anyone else, so the reference is run with the
library's masks substituted at the reference's own dropout call sites - the modules and their order stay the reference's:
  * nn.Dropout modules (visual/audio embedder[0], feature_encoding.dropout, layer.dropout1 / dropout / dropout2;
    helpers/encodings.py:141,149,177, helpers/transformers.py:74,79,82) get a forward hook that returns input * mask;
  * the attention-probability dropout inside F.multi_head_attention_forward (helpers/transformers.py:73: MultiheadAttention(dropout=p))
    is reached by putting only the MultiheadAttention modules in training mode and replacing torch.nn.functional.dropout while
    the model runs: its only training-mode callers are then the attention layers, in layer order.
The masks are oracle.tim_oracle_bwd.drop_mask (the numpy restatement of tim_b200/csrc/kernels.h: DropSite), converted to the
reference's tensor layouts ([B, S, E] before its transpose, [S, B, E] inside the layers, [B * H, S, S] for the probabilities).
Probabilities: the reference's defaults (feat_drop 0.5, seq_drop 0.5, enc_dropout 0.1; tim.py:22-28).
Fingerprints as in tools/make_golden_grads.py (L2 norm, sum, up to 512 seeded entries per gradient tensor).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.make_golden import CASES, INPUT_SEED, WEIGHT_SEED, build_reference   # noqa: E402
from tools.make_golden_grads import cotangent, fingerprint                       # noqa: E402

SEED = 0x5EEDD40F00D
P_FEAT, P_SEQ, P_ENC = 0.5, 0.5, 0.1
_small = dict(visual_input_dim=48, audio_input_dim=40, num_feats=6)
DROP_CASES = {
    "recog_av_small": CASES["recog_av_small"],
    "recog_visual": CASES["recog_visual"],
    "recog_audio": CASES["recog_audio"],
    "det_av": CASES["det_av"],
    "det_visual_vn": CASES["det_visual_vn"],
    # head widths the tcgen05 attention kernels serve (16-bit modes carry the probability dropout only there)
    "recog_hd64": (dict(num_class=[[5, 7, 11], 3], d_model=64, nhead=2, num_layers=2, **_small), 3, 5, 4),
    "det_hd128": (dict(num_class=(9, 4), d_model=128, nhead=2, num_layers=2, data_modality="audio_visual", include_verb_noun=False,
                       variant="detection", **_small), 2, 7, 7),
    "recog_hd64_wide": (dict(num_class=[[5, 7, 11], 3], d_model=64, nhead=2, num_layers=1, visual_input_dim=48, audio_input_dim=40,
                             num_feats=50), 2, 70, 70),
}


def install_masks(model, cfg, B, S):
    """Hooks + the functional patch described in the module docstring. Returns a context-manager-like (enter, exit) pair."""
    import torch
    import torch.nn.functional as F
    from oracle.tim_oracle_bwd import (DROP_FEAT_AUD, DROP_FEAT_VIS, DROP_FFN, DROP_SEQ, DROP_SUB1, DROP_SUB2, attn_drop_mask, drop_mask,
                                       row_drop_mask)
    Ft, E, H, FF = cfg.F_tot, 2 * cfg.d_model, cfg.nhead, cfg.FF
    handles = []

    def hook_with(mask_fn):
        def hook(_m, inputs, _out):
            x = inputs[0]
            m = torch.from_numpy(np.ascontiguousarray(mask_fn(tuple(x.shape)))).to(x.dtype)
            assert m.shape == x.shape, (m.shape, x.shape)
            return x * m
        return hook

    fe = model.feature_encoding
    for name, site in (("visual_embedder", DROP_FEAT_VIS), ("audio_embedder", DROP_FEAT_AUD)):
        emb = getattr(fe, name, None)
        if emb is not None:
            assert isinstance(emb[0], torch.nn.Dropout) and abs(emb[0].p - P_FEAT) < 1e-12
            handles.append(emb[0].register_forward_hook(
                hook_with(lambda shp, site=site: drop_mask(np.arange(int(np.prod(shp))).reshape(shp), P_FEAT, SEED, site))))
    assert abs(fe.dropout.p - P_SEQ) < 1e-12
    handles.append(fe.dropout.register_forward_hook(hook_with(lambda shp: row_drop_mask(B, S, Ft, E, P_SEQ, SEED, DROP_SEQ))))
    enc = getattr(model, cfg.encoder_prefix)
    for l, layer in enumerate(enc.layers):
        for mod, site, width in ((layer.dropout1, DROP_SUB1, E), (layer.dropout, DROP_FFN, FF), (layer.dropout2, DROP_SUB2, E)):
            assert abs(mod.p - P_ENC) < 1e-12
            handles.append(mod.register_forward_hook(                     # inside the layers the reference holds [S, B, C]
                hook_with(lambda shp, site=site, width=width, l=l: row_drop_mask(B, S, Ft, width, P_ENC, SEED, site, l).transpose(1, 0, 2))))
        assert abs(layer.self_attn.dropout - P_ENC) < 1e-12
        layer.self_attn.training = True
    state = {"layer": 0}
    real_dropout = F.dropout

    def attn_dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return real_dropout(x, p, training, inplace)
        l = state["layer"]
        state["layer"] += 1
        assert tuple(x.shape) == (B * H, S, S) and abs(p - P_ENC) < 1e-12, (tuple(x.shape), p)
        m = attn_drop_mask(B, H, S, Ft, P_ENC, SEED, l).reshape(B * H, S, S)
        return x * torch.from_numpy(np.ascontiguousarray(m)).to(x.dtype)

    F.dropout = attn_dropout

    def done():
        F.dropout = real_dropout
        for h in handles:
            h.remove()
        assert state["layer"] == cfg.num_layers, state
    return done


def run_variant(variant: str):
    import torch
    from tim_b200.config import TIMConfig
    from tim_b200.synth import synth_inputs, synth_state_dict
    torch.set_num_threads(8)
    blob = {}
    for name, (kw, B, Qv, Qa) in DROP_CASES.items():
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
        S = cfg.seq_len(Qv, Qa)
        done = install_masks(model, cfg, B, S)
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
        done()
        loss = 0.0
        for k, v in (("verb", verb), ("noun", noun), ("action", action), ("audio", audio), ("reg_v", reg_v), ("reg_a", reg_a),
                     ("feats", feats)):
            if v is not None:
                loss = loss + (v * torch.from_numpy(cotangent(name, k, v.shape))).sum()
                blob[f"{name}/out_shape/{k}"] = np.array(v.shape)
                blob[f"{name}/out/{k}"] = v.detach().numpy()
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
        print(f"[grads+dropout] {name}: {len(grads)} gradient tensors, loss {float(loss):.6f}", flush=True)
    return blob


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--variant":
        blob = run_variant(sys.argv[2])
        np.savez_compressed(sys.argv[3], **blob)
        return
    merged = {}
    for variant in ("recognition", "detection"):
        tmp = os.path.join(ROOT, "tests", "golden", f"_gradsd_{variant}.npz")
        r = subprocess.run([sys.executable, __file__, "--variant", variant, tmp], capture_output=True, text=True)
        sys.stdout.write(r.stdout)
        if r.returncode:
            sys.stderr.write(r.stderr)
            raise SystemExit(r.returncode)
        with np.load(tmp) as z:
            merged.update({k: z[k] for k in z.files})
        os.remove(tmp)
    merged["cases"] = np.array(list(DROP_CASES))
    merged["manifest"] = np.array(json.dumps({
        "seed": SEED, "p_feat": P_FEAT, "p_seq": P_SEQ, "p_enc": P_ENC, "weight_seed": WEIGHT_SEED, "input_seed": INPUT_SEED,
        "cases": {n: {"cfg": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}, "B": B, "Qv": Qv, "Qa": Qa}
                  for n, (kw, B, Qv, Qa) in DROP_CASES.items()}}))
    out = os.path.join(ROOT, "tests", "golden", "grads_dropout.npz")
    np.savez_compressed(out, **merged)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
