#!/usr/bin/env python
"""Mint golden vectors by running the UNMODIFIED reference (imported from /root/reference, CPU fp32).

The reference ships no numeric tests, so these fixtures are what pins the oracle
(oracle/tim_oracle.py) and, through it, the CUDA path. Weights and inputs are NOT stored: both are
regenerated bit-exactly from (seed, name) by tim_b200/synth.py; only the reference's outputs are.

Usage (in the build container, where /root/reference exists):
    python tools/make_golden.py            # regenerates tests/golden/*.npz + manifest.json
recognition and detection share the package name `time_interval_machine`, so each variant is
generated in its own subprocess.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (cfg kwargs, B, Qv, Qa)
CASES = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case, real widths
    "recog_cfg1": (dict(num_class=[[97, 300, 3806], 44], num_layers=1, num_feats=25), 2, 5, 5),
    # small-width cases covering every modality / head variant and the edge cases
    "recog_av_small": (dict(num_class=[[5, 7, 11], 3], visual_input_dim=48, audio_input_dim=40, d_model=64,
                            nhead=4, num_layers=2, num_feats=6), 3, 3, 2),
    "recog_av_novn": (dict(num_class=[11, 3], visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4,
                           num_layers=2, num_feats=6, include_verb_noun=False), 2, 4, 4),
    "recog_av_qa0": (dict(num_class=[[5, 7, 11], 3], visual_input_dim=48, audio_input_dim=40, d_model=64,
                          nhead=4, num_layers=1, num_feats=6), 2, 3, 0),
    "recog_av_datavis": (dict(num_class=[[5, 7, 11], 3], visual_input_dim=48, audio_input_dim=40, d_model=64,
                              nhead=4, num_layers=1, num_feats=6, data_modality="visual"), 2, 3, 0),
    "recog_visual": (dict(num_class=[[5, 7, 11], 3], visual_input_dim=48, d_model=64, nhead=4, num_layers=2,
                          num_feats=7, input_modality="visual", data_modality="visual"), 2, 3, 0),
    "recog_audio": (dict(num_class=[[5, 7, 11], 3], audio_input_dim=40, d_model=64, nhead=4, num_layers=2,
                         num_feats=7, input_modality="audio", data_modality="audio"), 2, 0, 4),
    "recog_hd192": (dict(num_class=[9, 4], visual_input_dim=64, audio_input_dim=64, d_model=96, nhead=1,
                         num_layers=1, num_feats=5, include_verb_noun=False), 1, 3, 3),
    "det_visual": (dict(num_class=[9, 4], visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4,
                        num_layers=2, num_feats=6, data_modality="visual", include_verb_noun=False,
                        variant="detection"), 2, 10, 0),
    "det_visual_vn": (dict(num_class=[[5, 7, 11], 4], visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4,
                           num_layers=1, num_feats=6, data_modality="visual", include_verb_noun=True,
                           variant="detection"), 2, 6, 0),
    "det_av": (dict(num_class=(9, 4), visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4,
                    num_layers=2, num_feats=6, data_modality="audio_visual", include_verb_noun=False,
                    variant="detection"), 2, 5, 5),
    "det_audio": (dict(num_class=[9, 4], visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4,
                       num_layers=1, num_feats=6, data_modality="audio", include_verb_noun=False,
                       variant="detection"), 2, 0, 5),
    # BASELINE.json configs[1] and configs[3] at their real widths and depth (6 layers), one clip each: fixtures minted by the
    # reference itself for the two configurations the headline numbers are quoted on
    "recog_cfg2": (dict(num_class=[[97, 300, 3806], 44], num_layers=6, num_feats=50), 1, 25, 25),
    "det_cfg4": (dict(num_class=[97, 44], visual_input_dim=2048, num_layers=6, num_feats=50, data_modality="visual",
                      include_verb_noun=False, variant="detection"), 1, 2048, 0),
    # the reference's own 399-query inference pyramid (detection/.../tim.py:140-155)
    "det_pyramid": (dict(num_class=[9, 4], visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4,
                         num_layers=1, num_feats=6, data_modality="visual", include_verb_noun=False,
                         variant="detection"), 1, -1, 0),
}
WEIGHT_SEED = 0
INPUT_SEED = 1234


def _install_shims():
    sj = types.ModuleType("simplejson")
    sj.dumps = json.dumps
    sys.modules["simplejson"] = sj
    fio = types.ModuleType("fvcore.common.file_io")
    fio.PathManager = type("PM", (), {"open": staticmethod(open)})
    sys.modules.update({"fvcore": types.ModuleType("fvcore"),
                        "fvcore.common": types.ModuleType("fvcore.common"),
                        "fvcore.common.file_io": fio})


def build_reference(cfg):
    """Instantiate the reference TIM for `cfg` (must run in a process dedicated to cfg.variant)."""
    import torch
    _install_shims()
    sys.path.insert(0, f"/root/reference/{cfg.variant}")
    from time_interval_machine.models.tim import TIM
    kw = dict(num_class=cfg.num_class, visual_input_dim=cfg.visual_input_dim, audio_input_dim=cfg.audio_input_dim,
              d_model=cfg.d_model, nhead=cfg.nhead, num_layers=cfg.num_layers, input_modality=cfg.input_modality,
              data_modality=cfg.data_modality, num_feats=cfg.num_feats, include_verb_noun=cfg.include_verb_noun)
    if cfg.variant == "recognition":
        kw["feedforward_scale"] = cfg.feedforward_scale
    else:
        kw["feedfoward_scale"] = cfg.feedforward_scale     # the reference's own spelling (detection tim.py:25)
    torch.manual_seed(0)
    return TIM(**kw).eval()


def run_variant(variant: str):
    import torch
    from tim_b200.config import TIMConfig, state_dict_spec
    from tim_b200.synth import synth_state_dict, synth_inputs, rel_l2
    from oracle.tim_oracle import TIMOracle

    torch.set_num_threads(8)
    manifest = {}
    only = set(filter(None, os.environ.get("GOLDEN_ONLY", "").split(",")))
    for name, (kw, B, Qv, Qa) in CASES.items():
        cfg = TIMConfig(**kw)
        if cfg.variant != variant or (only and name not in only):
            continue
        model = build_reference(cfg)
        ref_sd = model.state_dict()
        spec = state_dict_spec(cfg)
        assert list(ref_sd.keys()) == list(spec.keys()), (name, set(ref_sd) ^ set(spec))
        for k, v in ref_sd.items():
            assert tuple(v.shape) == tuple(spec[k]), (name, k, tuple(v.shape), spec[k])
        sd = synth_state_dict(cfg, WEIGHT_SEED, "trained")
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)

        pyramid = Qv < 0
        if pyramid:
            q = model.inference_queries.numpy().astype(np.float32)        # [1, 399, 2]
            Qv = q.shape[1]
        det = cfg.variant == "detection"
        inp = synth_inputs(cfg, B, Qv, Qa, INPUT_SEED, shared_queries=det)
        if pyramid:
            inp["times"][:, cfg.F_tot:] = q
        vis = torch.from_numpy(inp["vis"]) if "vis" in inp else None
        aud = torch.from_numpy(inp["aud"]) if "aud" in inp else None
        times = torch.from_numpy(inp["times"])
        out = {}
        with torch.no_grad():
            if not det:
                te = model(times, "time_mlp")
                (verb, noun, action, audio), feats = model([vis, aud], "encoder", te, Qv, Qa)
                out["time_encodings"] = te.numpy()
                reg_v = reg_a = None
            else:
                nq = max(Qv, Qa)
                model.inference_queries = times[0:1, cfg.F_tot:cfg.F_tot + nq].clone()
                model.num_queries = nq
                res = model([vis, aud], "encoder", times[:, :cfg.F_tot].clone(), None, False)
                (verb, noun, action, audio), (reg_v, reg_a), feats = res[0]
        for k, v in (("verb", verb), ("noun", noun), ("action", action), ("audio", audio),
                     ("reg_v", reg_v), ("reg_a", reg_a), ("feats", feats)):
            if v is not None:
                out[k] = v.numpy()

        # how close is the numpy restatement to the reference (fp32 vs fp32, and fp64 vs fp32)?
        errs = {}
        for dt in (np.float32, np.float64):
            o = TIMOracle(cfg, sd, dt).forward(inp.get("vis"), inp.get("aud"), inp["times"], Qv, Qa)
            errs[np.dtype(dt).name] = {k: rel_l2(o[k], v) for k, v in out.items()}
            assert all((o[k] is None) == (k not in out) for k in ("verb", "noun", "action", "audio", "reg_v", "reg_a"))
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        manifest[name] = dict(cfg=kw, B=B, Qv=Qv, Qa=Qa, weight_seed=WEIGHT_SEED, input_seed=INPUT_SEED,
                              style="trained", shared_queries=det, pyramid=pyramid,
                              outputs={k: list(v.shape) for k, v in out.items()}, oracle_rel_l2=errs)
        worst = max(max(e.values()) for e in errs.values())
        print(f"[golden] {name}: {len(out)} tensors, oracle-vs-reference worst rel-L2 {worst:.2e}", flush=True)
    return manifest


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--variant":
        m = run_variant(sys.argv[2])
        print("MANIFEST " + json.dumps(m))
        return
    os.makedirs(GOLD, exist_ok=True)
    manifest = {"reference_commit": "c7cb2935fb6736db3e4328e749ad25a2433a1654", "generator": "tools/make_golden.py",
                "cases": {}}
    # GOLDEN_ONLY=a,b mints only those cases and merges them into the existing manifest (the other .npz files stay byte-identical)
    mp = os.path.join(GOLD, "manifest.json")
    if os.environ.get("GOLDEN_ONLY") and os.path.exists(mp):
        manifest = json.load(open(mp))
    for variant in ("recognition", "detection"):
        r = subprocess.run([sys.executable, __file__, "--variant", variant], capture_output=True, text=True)
        sys.stdout.write("\n".join(l for l in r.stdout.splitlines() if not l.startswith("MANIFEST ")) + "\n")
        if r.returncode:
            sys.stderr.write(r.stderr)
            raise SystemExit(r.returncode)
        line = [l for l in r.stdout.splitlines() if l.startswith("MANIFEST ")][-1]
        manifest["cases"].update(json.loads(line[len("MANIFEST "):]))
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(f"wrote {len(manifest['cases'])} cases to {GOLD}")


if __name__ == "__main__":
    main()
