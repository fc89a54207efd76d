"""One process per reference variant (both use the package name `time_interval_machine`): the UNMODIFIED reference TIM on cuda:0,
un-patched (eager PyTorch, fp32, TF32 off) against the same module object after tim_b200.patch_model().

    python tests/_ref_worker.py <recognition|detection> <eval|train> <out.json>

eval : forward parity of the patched module (fp32 and fp16 compute) against its own un-patched forward on the same GPU.
train: gradient parity - the patched module in train() mode (dropout p = 0), loss = sum <output, cotangent>, backward through
       tim_b200's autograd.Function, against torch.autograd over the un-patched module (fp32, TF32 off) on the same GPU.
Test infrastructure only (imports the reference through tools/refload.py).
"""
import json
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tim_b200.config import TIMConfig, named_config   # noqa: E402
from tim_b200.synth import synth_inputs, synth_state_dict   # noqa: E402
from tools.refload import build_reference   # noqa: E402

# (name, config kwargs or named config, B, Qv, Qa)
CASES = {
    "recognition": [
        ("recog_small", dict(num_class=[[5, 7, 11], 3], visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4,
                             num_layers=2, num_feats=6), 3, 4, 3),
        ("cfg1", "cfg1", 2, 5, 5),
        ("cfg2", "cfg2", 2, 25, 25),
    ],
    "detection": [
        ("det_small", dict(num_class=[9, 4], visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4, num_layers=2,
                           num_feats=6, data_modality="visual", include_verb_noun=False, variant="detection"), 2, 10, 0),
        ("det_av", dict(num_class=(9, 4), visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4, num_layers=2, num_feats=6,
                        data_modality="audio_visual", include_verb_noun=False, variant="detection"), 2, 5, 5),
        ("cfg4", "cfg4", 1, 2048, 0),
    ],
}


def rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)


def worst_row_rel(a, b):
    """max over rows of ||a_r - b_r|| / ||b_r - mean(b_r)|| (the per-row check ADVICE.md asks for: a constant logit bias does not
    hide errors of individual query rows)."""
    a, b = a.detach().double(), b.detach().double()
    a, b = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
    den = (b - b.mean(dim=1, keepdim=True)).norm(dim=1).clamp_min(1e-12)
    return float(((a - b).norm(dim=1) / den).max())


def cotangent(case, key, shape, dev):
    rng = np.random.default_rng(zlib.crc32(f"{case}/{key}".encode()))
    return torch.from_numpy(rng.standard_normal(tuple(shape)).astype(np.float32)).to(dev)


def set_dropout_zero(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0


def synth_target(cfg, B, dev, n_gt=3):
    """A loader-style target dict for detection forward_train (detection/.../models/tim.py:214-222 reads these keys)."""
    g = torch.Generator().manual_seed(4242)
    st = torch.rand((B, n_gt), generator=g) * 0.8
    seg = torch.stack([st, st + 0.05 + 0.3 * torch.rand((B, n_gt), generator=g)], dim=-1)
    hc = cfg.head_classes()
    tgt = {"v_gt_segments": seg.to(dev), "a_gt_segments": seg.flip(1).contiguous().to(dev)}
    tgt["verb"] = torch.randint(0, max(hc["verb"], 1), (B, n_gt), generator=g).to(dev)
    tgt["noun"] = torch.randint(0, max(hc["noun"], 1), (B, n_gt), generator=g).to(dev)
    tgt["action"] = torch.randint(0, max(hc["action"], 1), (B, n_gt), generator=g).to(dev)
    tgt["class_id"] = torch.randint(0, max(hc["audio"], 1), (B, n_gt), generator=g).to(dev)
    return tgt


def run_model(model, cfg, t, Qv, Qa, train=False):
    """-> dict of output tensors (None where absent), through the reference's own call signatures."""
    if cfg.variant == "recognition":
        te = model(t["times"], "time_mlp")
        (verb, noun, action, audio), feats = model([t.get("vis"), t.get("aud")], "encoder", te, Qv, Qa)
        return dict(verb=verb, noun=noun, action=action, audio=audio, feats=feats)
    nq = max(Qv, Qa)
    pool = t["times"][0:1, cfg.F_tot:cfg.F_tot + nq].clone()
    model.num_queries = nq
    if train:
        # forward_train draws its queries from model.train_pool with the CPU global RNG (tim.py:281-282): same seed, same draw
        model.train_pool = pool.cpu()
        torch.manual_seed(777)
        B = t["times"].shape[0]
        res = model([t.get("vis"), t.get("aud")], "encoder", t["times"][:, :cfg.F_tot].clone(), synth_target(cfg, B, pool.device))
    else:
        model.inference_queries = pool
        res = model([t.get("vis"), t.get("aud")], "encoder", t["times"][:, :cfg.F_tot].clone(), None, False)
    (verb, noun, action, audio), (reg_v, reg_a), feats = res[0]
    return dict(verb=verb, noun=noun, action=action, audio=audio, reg_v=reg_v, reg_a=reg_a, feats=feats)


def main():
    variant, mode, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    from tim_b200.plugin import patch_model
    report = {"variant": variant, "mode": mode, "cases": {}}
    for name, spec, B, Qv, Qa in CASES[variant]:
        cfg = named_config(spec)[0] if isinstance(spec, str) else TIMConfig(**spec)
        sd = synth_state_dict(cfg, 0, "trained")
        inp = synth_inputs(cfg, B, Qv, Qa, 1234, shared_queries=cfg.variant == "detection")
        t = {k: torch.from_numpy(v).to(dev) for k, v in inp.items()}
        model = build_reference(cfg).to(dev)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        rec = {}
        if mode == "eval":
            model.eval()
            torch.manual_seed(4321)
            with torch.no_grad():
                ref = run_model(model, cfg, t, Qv, Qa)
            rng_ref = torch.get_rng_state()        # the reference's detection inference draws from the CPU generator (tim.py:364)
            ref = {k: (v.clone() if v is not None else None) for k, v in ref.items()}
            for dt in ("fp32", "fp16"):
                m2 = build_reference(cfg).to(dev)
                m2.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
                m2 = patch_model(m2.eval(), compute_dtype=dt)
                torch.manual_seed(4321)
                with torch.no_grad():
                    got = run_model(m2, cfg, t, Qv, Qa)
                torch.cuda.synchronize()
                # the drop-in leaves torch's CPU generator where the reference leaves it: whatever the caller draws next is unchanged
                assert torch.equal(torch.get_rng_state(), rng_ref), (name, dt, "CPU RNG state differs from the reference's after the forward")
                for k, v in ref.items():
                    assert (v is None) == (got[k] is None), (name, k)
                    if v is not None:
                        assert tuple(v.shape) == tuple(got[k].shape), (name, k, v.shape, got[k].shape)
                        rec[f"{dt}/{k}"] = {"rel_l2": rel(got[k], v), "worst_row": worst_row_rel(got[k], v)}
                del m2
        else:
            set_dropout_zero(model)
            model.train()
            out = run_model(model, cfg, t, Qv, Qa, train=True)
            loss = sum((v * cotangent(name, k, v.shape, dev)).sum() for k, v in out.items() if v is not None)
            loss.backward()
            ref_g = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
            # yardstick: the reference's OWN mixed precision (fp16 autocast, as its training loop runs it, train.py:197) against
            # its fp32 run - gradients behind a ReLU gate move by the same order in any 16-bit implementation
            model.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.float16):
                out_ac = run_model(model, cfg, t, Qv, Qa, train=True)
                loss_ac = sum((v.float() * cotangent(name, k, v.shape, dev)).sum() for k, v in out_ac.items() if v is not None)
            loss_ac.backward()
            for k, g in ref_g.items():
                pg = dict(model.named_parameters())[k].grad
                if pg is not None and not (k.startswith("drloc_mlp") or k.startswith("pool")):
                    rec[f"ref_autocast_fp16/{k}"] = {"rel_l2": rel(pg, g)}
            model.zero_grad(set_to_none=True)
            for dt in ("fp32", "fp16"):
                m2 = build_reference(cfg).to(dev)
                m2.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
                set_dropout_zero(m2)
                m2 = patch_model(m2.train(), compute_dtype=dt)
                out2 = run_model(m2, cfg, t, Qv, Qa, train=True)
                loss2 = sum((v * cotangent(name, k, v.shape, dev)).sum() for k, v in out2.items() if v is not None)
                loss2.backward()
                torch.cuda.synchronize()
                got_g = {k: p.grad for k, p in m2.named_parameters() if p.grad is not None}
                for k, g in ref_g.items():
                    if k.startswith("drloc_mlp") or k.startswith("pool"):
                        continue
                    assert k in got_g, (name, dt, "missing gradient", k)
                    rec[f"{dt}/{k}"] = {"rel_l2": rel(got_g[k], g)}
                rec[f"{dt}/loss"] = {"rel_l2": abs(float(loss2) - float(loss)) / max(abs(float(loss)), 1e-12)}
                del m2
        report["cases"][name] = rec
        del model
        torch.cuda.empty_cache()
    json.dump(report, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
