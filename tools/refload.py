"""Imports the UNMODIFIED reference TIM class (test / measurement infrastructure only; nothing under tim_b200/ uses it).

Looks for the reference package at /root/reference/<variant> (build container) and at baseline/_ref/<variant> (the staged copy that
travels to the GPU box, tools/stage_reference.py). Both variants use the package name `time_interval_machine`, so ONE process can
import only one of them: callers run one process per variant (tests/_ref_worker.py).
Two logging-only third-party modules the reference imports are absent from this image and are shimmed in memory.
"""
from __future__ import annotations

import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root(variant: str):
    for base in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        p = os.path.join(base, variant)
        if os.path.isfile(os.path.join(p, "time_interval_machine", "models", "tim.py")):
            return p
    return None


def install_shims():
    if "simplejson" not in sys.modules:
        sj = types.ModuleType("simplejson")
        sj.dumps = json.dumps
        sys.modules["simplejson"] = sj
    if "fvcore.common.file_io" not in sys.modules:
        fio = types.ModuleType("fvcore.common.file_io")
        fio.PathManager = type("PM", (), {"open": staticmethod(open)})
        sys.modules.update({"fvcore": types.ModuleType("fvcore"), "fvcore.common": types.ModuleType("fvcore.common"),
                            "fvcore.common.file_io": fio})


def build_reference(cfg, seed: int = 0):
    """Instantiate the reference TIM for `cfg` (CPU, eval mode). Raises FileNotFoundError when no copy of the reference exists."""
    import torch
    root = reference_root(cfg.variant)
    if root is None:
        raise FileNotFoundError(f"reference package for '{cfg.variant}' not found (neither /root/reference nor baseline/_ref)")
    loaded = sys.modules.get("time_interval_machine")
    if loaded is not None and not os.path.abspath(loaded.__file__).startswith(os.path.abspath(root)):
        raise RuntimeError("another variant of time_interval_machine is already imported in this process")
    install_shims()
    if root not in sys.path:
        sys.path.insert(0, root)
    from time_interval_machine.models.tim import TIM
    kw = dict(num_class=cfg.num_class, visual_input_dim=cfg.visual_input_dim, audio_input_dim=cfg.audio_input_dim,
              d_model=cfg.d_model, nhead=cfg.nhead, num_layers=cfg.num_layers, input_modality=cfg.input_modality,
              data_modality=cfg.data_modality, num_feats=cfg.num_feats, include_verb_noun=cfg.include_verb_noun)
    if cfg.variant == "recognition":
        kw["feedforward_scale"] = cfg.feedforward_scale
    else:
        kw["feedfoward_scale"] = cfg.feedforward_scale     # the reference's own spelling (detection tim.py:25)
    torch.manual_seed(seed)
    return TIM(**kw).eval()
