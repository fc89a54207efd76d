"""Host side of the drop-in: the reference's model interface on top of libtim_b200.

The reference exposes no plugin API; its seam is TIM.forward (recognition/.../models/tim.py:174-191,
detection/.../models/tim.py:415-430) and the state_dict layout. This module gives
  * TIMEngine        - one library context: load_state_dict / time_mlp / encoder / forward_host
  * patch_model()    - rebinds forward() of an existing reference TIM instance (same signature, same return
                       structure, parameters stay nn.Parameters under the same names, so the reference's
                       train/eval scripts, flags and checkpoints are untouched)
PyTorch is used for device memory and streams only. There is no fallback: every call goes through the
C ABI and raises if the library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import types
from typing import Dict, Mapping, Optional

import numpy as np
import torch

from . import _lib
from .config import (DETECTION, DTYPE_CODES, RECOGNITION, TIMConfig, hot_path_keys, state_dict_spec)


def _c_config(cfg: TIMConfig, compute_dtype: str) -> _lib.tim_config:
    hc = cfg.head_classes()
    return _lib.tim_config(
        variant=_lib.VARIANT_CODES[cfg.variant], d_model=cfg.d_model, nhead=cfg.nhead, num_layers=cfg.num_layers,
        ff_dim=cfg.FF, vis_dim=cfg.visual_input_dim, aud_dim=cfg.audio_input_dim, num_feats=cfg.num_feats,
        input_modality=_lib.MODALITY_CODES[cfg.input_modality], data_modality=_lib.MODALITY_CODES[cfg.data_modality],
        include_verb_noun=int(cfg.verb_noun_tokens), n_verb=hc["verb"], n_noun=hc["noun"], n_action=hc["action"],
        n_audio=hc["audio"], compute_dtype=DTYPE_CODES[compute_dtype])


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class TIMEngine:
    """One tim_ctx on one CUDA device."""

    def __init__(self, cfg: TIMConfig, device: int = 0, compute_dtype: str = "fp16"):
        if compute_dtype not in DTYPE_CODES:
            raise ValueError(f"compute_dtype must be one of {sorted(DTYPE_CODES)}")
        self.lib = _lib.load()
        self.cfg = cfg
        self.compute_dtype = compute_dtype
        self.device = torch.device("cuda", device)
        self._ctx = C.c_void_p()
        self._cc = _c_config(cfg, compute_dtype)
        _lib.check(self.lib.tim_create(C.byref(self._ctx), C.byref(self._cc), device))
        self._keys = hot_path_keys(cfg)
        self._spec = state_dict_spec(cfg, include_drloc=False)

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.tim_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_weight(self, key: str, value) -> None:
        if isinstance(value, np.ndarray):
            value = torch.from_numpy(value)
        t = value.detach().to(device=self.device, dtype=torch.float32).contiguous()
        shape = (C.c_int64 * t.dim())(*t.shape)
        _lib.check(self.lib.tim_set_weight(self._ctx, key.encode(), _ptr(t), shape, t.dim(), self._stream()), self._ctx)
        # the pack kernels were enqueued on the current stream; t may be a temporary
        t.record_stream(torch.cuda.current_stream(self.device))

    def load_state_dict(self, sd: Mapping[str, object], strict: bool = True) -> None:
        """Accepts the reference's state_dict (extra keys such as drloc_mlp.* are ignored)."""
        with torch.cuda.device(self.device):
            for k in self._keys:
                if k in sd:
                    self.set_weight(k, sd[k])
                elif strict:
                    raise KeyError(f"state_dict is missing '{k}'")

    def missing_weights(self):
        buf = C.create_string_buffer(1 << 16)
        n = self.lib.tim_weights_missing(self._ctx, buf, len(buf))
        return [s for s in buf.value.decode().split("\n") if s] if n > 0 else []

    # ------------------------------------------------------------------ forward (device tensors)
    def _check_in(self, t: torch.Tensor, name: str, shape):
        if t.device != self.device:
            raise ValueError(f"{name} must live on {self.device}, got {t.device}")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
        return t.to(torch.float32).contiguous()

    def time_mlp(self, times: torch.Tensor) -> torch.Tensor:
        """tim.py:181-182 — times [B, T, 2] -> [B, T, d_model]."""
        B, T = int(times.shape[0]), int(times.shape[1])
        times = self._check_in(times, "times", (B, T, 2))
        with torch.cuda.device(self.device):
            out = torch.empty((B, T, self.cfg.d_model), device=self.device, dtype=torch.float32)
            _lib.check(self.lib.tim_time_mlp_fwd(self._ctx, _ptr(times), _ptr(out), B, T, self._stream()), self._ctx)
        return out

    def _alloc_outputs(self, B: int, Qv: int, Qa: int, *, pinned: bool, want_feats: bool = True, dtype=torch.float32):
        cfg = self.cfg
        hc = cfg.head_classes()
        qv = Qv if "visual" in cfg.data_modality else 0
        qa = Qa if "audio" in cfg.data_modality else 0
        kw = dict(dtype=dtype, device="cpu", pin_memory=True) if pinned else dict(dtype=dtype, device=self.device)
        out: Dict[str, Optional[torch.Tensor]] = {k: None for k in ("verb", "noun", "action", "audio", "reg_v", "reg_a", "feats")}
        if "visual" in cfg.data_modality:
            if hc["verb"]:
                out["verb"] = torch.empty((B * qv, hc["verb"]), **kw)
                out["noun"] = torch.empty((B * qv, hc["noun"]), **kw)
            out["action"] = torch.empty((B * qv, hc["action"]), **kw)
            if cfg.has_reg_head:
                out["reg_v"] = torch.empty((B * qv, 2), **kw)
        if "audio" in cfg.data_modality:
            out["audio"] = torch.empty((B * qa, hc["audio"]), **kw)
            if cfg.has_reg_head:
                out["reg_a"] = torch.empty((B * qa, 2), **kw)
        if want_feats:
            out["feats"] = torch.empty((B, cfg.F_tot, cfg.E), **kw)
        co = _lib.tim_outputs(verb=_ptr(out["verb"]), noun=_ptr(out["noun"]), action=_ptr(out["action"]),
                              audio=_ptr(out["audio"]), reg_visual=_ptr(out["reg_v"]), reg_audio=_ptr(out["reg_a"]),
                              feats=_ptr(out["feats"]))
        return out, co

    def encoder(self, vis: Optional[torch.Tensor], aud: Optional[torch.Tensor], time_enc: torch.Tensor,
                Qv: int, Qa: int, want_feats: bool = True) -> Dict[str, Optional[torch.Tensor]]:
        """tim.py:147-172 — returns dict(verb, noun, action, audio, reg_v, reg_a, feats); None where absent."""
        cfg = self.cfg
        B, T = int(time_enc.shape[0]), int(time_enc.shape[1])
        Qv, Qa = int(Qv or 0), int(Qa or 0)
        time_enc = self._check_in(time_enc, "time_encodings", (B, T, cfg.d_model))
        if cfg.has_visual_input:
            vis = self._check_in(vis, "visual input", (B, cfg.num_feats, cfg.visual_input_dim))
        if cfg.has_audio_input:
            aud = self._check_in(aud, "audio input", (B, cfg.num_feats, cfg.audio_input_dim))
        with torch.cuda.device(self.device):
            out, co = self._alloc_outputs(B, Qv, Qa, pinned=False, want_feats=want_feats)
            _lib.check(self.lib.tim_encoder_fwd(self._ctx, _ptr(vis if cfg.has_visual_input else None),
                                                _ptr(aud if cfg.has_audio_input else None), _ptr(time_enc),
                                                B, T, Qv, Qa, C.byref(co), self._stream()), self._ctx)
        return out

    def encoder_indexed(self, vis_bank: Optional[torch.Tensor], vis_rows: Optional[torch.Tensor], aud_bank: Optional[torch.Tensor],
                        aud_rows: Optional[torch.Tensor], time_enc: torch.Tensor, Qv: int, Qa: int,
                        want_feats: bool = True) -> Dict[str, Optional[torch.Tensor]]:
        """encoder() with the input windows gathered on the device from feature banks resident in HBM: *_bank [rows, dim] (fp32,
        fp16 or bf16, same dtype for both), *_rows [B, num_feats] int64 = bank row of every feature token (what the reference's
        loader indexes on the host: datasets/sliding_window.py:356-375). A row index outside its bank is an error, reported
        asynchronously: by index_check() (blocks) or by the next encoder_indexed() on this engine."""
        cfg = self.cfg
        B, T = int(time_enc.shape[0]), int(time_enc.shape[1])
        time_enc = self._check_in(time_enc, "time_encodings", (B, T, cfg.d_model))
        codes = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
        fb = _lib.tim_feature_bank()
        dt = None
        for name, bank, rows, dim, present in (("vis", vis_bank, vis_rows, cfg.visual_input_dim, cfg.has_visual_input),
                                               ("aud", aud_bank, aud_rows, cfg.audio_input_dim, cfg.has_audio_input)):
            if not present:
                continue
            if bank is None or rows is None or bank.device != self.device or rows.device != self.device:
                raise ValueError(f"{name}: bank and row indices must be tensors on {self.device}")
            if bank.dim() != 2 or bank.shape[1] != dim or not bank.is_contiguous() or bank.dtype not in codes:
                raise ValueError(f"{name}_bank must be contiguous [rows, {dim}] fp32 / fp16 / bf16")
            if rows.dtype != torch.int64 or tuple(rows.shape) != (B, cfg.num_feats) or not rows.is_contiguous():
                raise ValueError(f"{name}_rows must be contiguous int64 [{B}, {cfg.num_feats}]")
            if dt is not None and bank.dtype != dt:
                raise ValueError("both banks must have the same dtype")
            dt = bank.dtype
            setattr(fb, f"{name}_bank", bank.data_ptr()); setattr(fb, f"{name}_rows", rows.data_ptr())
            setattr(fb, f"{name}_bank_rows", int(bank.shape[0]))
        fb.bank_dtype = codes[dt]
        with torch.cuda.device(self.device):
            out, co = self._alloc_outputs(B, int(Qv or 0), int(Qa or 0), pinned=False, want_feats=want_feats)
            _lib.check(self.lib.tim_encoder_fwd_indexed(self._ctx, C.byref(fb), _ptr(time_enc), B, T, int(Qv or 0), int(Qa or 0),
                                                        C.byref(co), self._stream()), self._ctx)
        return out

    # ------------------------------------------------------------------ training leg
    def enable_training(self) -> None:
        """Allocates the transposed weight copies of the input-gradient GEMMs; weights must be (re-)set afterwards."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tim_train_enable(self._ctx), self._ctx)

    def set_dropout(self, p_feat: float = 0.0, p_seq: float = 0.0, p_enc: float = 0.0, seed: int = 0) -> None:
        """Dropout of the next encoder_train (and its backward): the reference's feat_drop / seq_drop / enc_dropout (tim.py:22-28).
        Masks are a counter-based hash of (seed, site, layer, element) - pass a fresh seed per step; all zero = no dropout."""
        _lib.check(self.lib.tim_set_dropout(self._ctx, float(p_feat), float(p_seq), float(p_enc), int(seed) & 0xFFFFFFFFFFFFFFFF), self._ctx)

    def bind_grad(self, key: str, grad: torch.Tensor) -> None:
        """The backward accumulates d loss / d <key> into `grad` (fp32, contiguous, on this device)."""
        if grad.device != self.device or grad.dtype != torch.float32 or not grad.is_contiguous():
            raise ValueError(f"gradient buffer of '{key}' must be a contiguous fp32 tensor on {self.device}")
        _lib.check(self.lib.tim_bind_grad(self._ctx, key.encode(), _ptr(grad)), self._ctx)

    def time_mlp_train(self, times: torch.Tensor) -> torch.Tensor:
        B, T = int(times.shape[0]), int(times.shape[1])
        times = self._check_in(times, "times", (B, T, 2))
        with torch.cuda.device(self.device):
            out = torch.empty((B, T, self.cfg.d_model), device=self.device, dtype=torch.float32)
            _lib.check(self.lib.tim_time_mlp_fwd_train(self._ctx, _ptr(times), _ptr(out), B, T, self._stream()), self._ctx)
        return out

    def time_mlp_bwd(self, d_out: torch.Tensor) -> None:
        d_out = d_out.to(device=self.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tim_time_mlp_bwd(self._ctx, _ptr(d_out), self._stream()), self._ctx)
        d_out.record_stream(torch.cuda.current_stream(self.device))

    def encoder_train(self, vis, aud, time_enc: torch.Tensor, Qv: int, Qa: int) -> Dict[str, Optional[torch.Tensor]]:
        cfg = self.cfg
        B, T = int(time_enc.shape[0]), int(time_enc.shape[1])
        Qv, Qa = int(Qv or 0), int(Qa or 0)
        time_enc = self._check_in(time_enc, "time_encodings", (B, T, cfg.d_model))
        if cfg.has_visual_input:
            vis = self._check_in(vis, "visual input", (B, cfg.num_feats, cfg.visual_input_dim))
        if cfg.has_audio_input:
            aud = self._check_in(aud, "audio input", (B, cfg.num_feats, cfg.audio_input_dim))
        with torch.cuda.device(self.device):
            out, co = self._alloc_outputs(B, Qv, Qa, pinned=False, want_feats=True)
            _lib.check(self.lib.tim_encoder_fwd_train(self._ctx, _ptr(vis if cfg.has_visual_input else None),
                                                      _ptr(aud if cfg.has_audio_input else None), _ptr(time_enc),
                                                      B, T, Qv, Qa, C.byref(co), self._stream()), self._ctx)
        self._train_shape = (B, T)
        return out

    def encoder_bwd(self, grads: Mapping[str, Optional[torch.Tensor]]) -> torch.Tensor:
        """grads: gradients of the encoder_train outputs (None where the loss does not reach one). Returns d loss / d time_enc."""
        B, T = self._train_shape
        keep = {}
        for k in ("verb", "noun", "action", "audio", "reg_v", "reg_a", "feats"):
            g = grads.get(k)
            keep[k] = g.to(device=self.device, dtype=torch.float32).contiguous() if g is not None else None
        go = _lib.tim_outputs(verb=_ptr(keep["verb"]), noun=_ptr(keep["noun"]), action=_ptr(keep["action"]), audio=_ptr(keep["audio"]),
                              reg_visual=_ptr(keep["reg_v"]), reg_audio=_ptr(keep["reg_a"]), feats=_ptr(keep["feats"]))
        with torch.cuda.device(self.device):
            d_te = torch.empty((B, T, self.cfg.d_model), device=self.device, dtype=torch.float32)
            _lib.check(self.lib.tim_encoder_bwd(self._ctx, C.byref(go), _ptr(d_te), self._stream()), self._ctx)
            st = torch.cuda.current_stream(self.device)
            for g in keep.values():
                if g is not None:
                    g.record_stream(st)
        return d_te

    def comm_init(self, group=None) -> None:
        """One NCCL communicator of the library's own over the ranks of `group` (the id travels through torch.distributed)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = (C.c_char * 128)()
        if rank == 0:
            _lib.check(self.lib.tim_comm_unique_id(buf), None)
        obj = [bytes(buf) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idbuf = C.create_string_buffer(obj[0], 128)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tim_comm_init(self._ctx, idbuf, rank, world), self._ctx)
        self._comm_world = world

    def allreduce(self, flat: torch.Tensor) -> None:
        """The one data-path collective of a training step: average `flat` (fp32, contiguous) over the ranks, in place."""
        if flat.device != self.device or flat.dtype != torch.float32 or not flat.is_contiguous():
            raise ValueError("allreduce needs a contiguous fp32 tensor on the engine's device")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tim_allreduce_grads(self._ctx, _ptr(flat), flat.numel(), self._stream()), self._ctx)

    @property
    def tape_bytes(self) -> int:
        return int(self.lib.tim_train_tape_bytes(self._ctx))

    # ------------------------------------------------------------------ forward (host tensors, end to end)
    def forward_host(self, vis, aud, times: torch.Tensor, Qv: int, Qa: int, clips_per_chunk: int = 0,
                     want_feats: bool = True, out=None, out_dtype=torch.float32):
        """time_mlp + encoder on HOST tensors (pinned for full copy bandwidth); H2D/D2H copies are inside the call.
        vis / aud: fp32, or the engine's 16-bit operand dtype (a 16-bit host feature bank: bit-identical results on the 16-bit
        paths, half the H2D bytes, no cast pass). out_dtype: torch.float32 or torch.float16 (logits rounded once on the device).
        Returns (outputs dict of pinned host tensors, h2d_bytes, d2h_bytes)."""
        cfg = self.cfg
        B, T = int(times.shape[0]), int(times.shape[1])
        op = {"fp16": torch.float16, "bf16": torch.bfloat16}.get(self.compute_dtype)
        feat_dt = None
        for name, t in (("vis", vis), ("aud", aud)):
            if t is None:
                continue
            if t.device.type != "cpu" or not t.is_contiguous() or t.dtype not in (torch.float32, op):
                raise ValueError(f"{name} must be a contiguous CPU tensor, fp32 or the engine's 16-bit operand dtype")
            if feat_dt is not None and t.dtype != feat_dt:
                raise ValueError("vis and aud must have the same dtype")
            feat_dt = t.dtype
        if times.device.type != "cpu" or times.dtype != torch.float32 or not times.is_contiguous():
            raise ValueError("times must be a contiguous fp32 CPU tensor")
        if out_dtype not in (torch.float32, torch.float16):
            raise ValueError("out_dtype must be torch.float32 or torch.float16")
        if out is None:
            out = self._alloc_outputs(B, Qv, Qa, pinned=True, want_feats=want_feats, dtype=out_dtype)
        outs, co = out
        for v in outs.values():
            if v is not None and v.dtype != out_dtype:
                raise ValueError("preallocated outputs do not match out_dtype")
        in_code = DTYPE_CODES["fp32"] if feat_dt in (None, torch.float32) else DTYPE_CODES[self.compute_dtype]
        out_code = DTYPE_CODES["fp32"] if out_dtype == torch.float32 else DTYPE_CODES["fp16"]
        up, down = C.c_uint64(0), C.c_uint64(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tim_forward_host_ex(self._ctx, _ptr(vis if cfg.has_visual_input else None),
                                                    _ptr(aud if cfg.has_audio_input else None), _ptr(times), B, T,
                                                    int(Qv or 0), int(Qa or 0), C.byref(co), int(clips_per_chunk), in_code, out_code,
                                                    self._stream(), C.byref(up), C.byref(down)), self._ctx)
        return outs, int(up.value), int(down.value)

    def index_check(self) -> None:
        """Blocks until the row indices of the last encoder_indexed() are known to have been inside their banks; raises TimError
        otherwise (the reference's host-side indexing raises IndexError; see tim_index_check in include/tim_b200.h)."""
        _lib.check(self.lib.tim_index_check(self._ctx), self._ctx)

    def fold_check(self) -> bool:
        """Blocks until the last encoder forward's precision check is known; True = recompute it (see tim_fold_check)."""
        return bool(_lib.check_nonneg(self.lib.tim_fold_check(self._ctx), self._ctx))

    # ------------------------------------------------------------------ detection query labelling
    def label_queries(self, queries: torch.Tensor, gt_segs: torch.Tensor, gt_labels: torch.Tensor, iou_threshold: float):
        """detection/.../models/tim.py:186-270 on the device: (targets [B*Nq,2], label ids [B*Nq,Nl] (-1 = negative), ious [B*Nq])."""
        q = queries.to(device=self.device, dtype=torch.float32).contiguous()
        g = gt_segs.to(device=self.device, dtype=torch.float32).contiguous()
        lab = gt_labels.to(device=self.device, dtype=torch.int64).contiguous()
        B, Nq, Na, Nl = int(q.shape[0]), int(q.shape[1]), int(g.shape[1]), int(lab.shape[2])
        if q.shape != (B, Nq, 2) or g.shape != (B, Na, 2) or lab.shape != (B, Na, Nl):
            raise ValueError("label_queries: queries [B,Nq,2], gt_segs [B,Na,2], gt_labels [B,Na,Nl] expected")
        with torch.cuda.device(self.device):
            targets = torch.empty((B * Nq, 2), dtype=torch.float32, device=self.device)
            ids = torch.empty((B * Nq, Nl), dtype=torch.int64, device=self.device)
            ious = torch.empty((B * Nq,), dtype=torch.float32, device=self.device)
            _lib.check(self.lib.tim_label_queries(_ptr(q), _ptr(g), _ptr(lab), B, Nq, Na, Nl, float(iou_threshold), _ptr(targets),
                                                  _ptr(ids), _ptr(ious), self._stream()), None)
        return targets, ids, ious

    def smooth_labels(self, ids: torch.Tensor, col: int, num_classes: int, smoothing: float) -> torch.Tensor:
        """assign_positive_labels (tim.py:158-185) for one label column: [rows, num_classes] fp32."""
        rows, stride = int(ids.shape[0]), int(ids.shape[1])
        with torch.cuda.device(self.device):
            out = torch.empty((rows, num_classes), dtype=torch.float32, device=self.device)
            if rows:
                _lib.check(self.lib.tim_smooth_labels(_ptr(ids), stride, int(col), rows, int(num_classes), float(smoothing), _ptr(out),
                                                      self._stream()), None)
        return out

    PROFILE_CLASSES = ("gemm_other", "attention", "layernorm", "assemble", "other", "gemm_in_proj_linear1", "gemm_out_proj_linear2",
                       "gemm_dgrad", "gemm_wgrad", "attention_bwd")

    def profile_begin(self) -> None:
        _lib.check(self.lib.tim_profile_begin(self._ctx), self._ctx)

    def profile_end(self) -> Dict[str, Dict[str, float]]:
        """{class: {ms, flops, launches}} accumulated since profile_begin() (synchronises the device)."""
        n = len(self.PROFILE_CLASSES)
        ms, fl, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_uint64 * n)()
        _lib.check(self.lib.tim_profile_end(self._ctx, ms, fl, cnt, n), self._ctx)
        out = {name: {"ms": ms[i], "flops": fl[i], "launches": int(cnt[i])} for i, name in enumerate(self.PROFILE_CLASSES)}
        # "gemm" = every dense contraction of the step (the three GEMM classes together)
        parts = [out[k] for k in ("gemm_other", "gemm_in_proj_linear1", "gemm_out_proj_linear2", "gemm_dgrad", "gemm_wgrad")]
        out["gemm"] = {k: sum(p[k] for p in parts) for k in ("ms", "flops", "launches")}
        return out

    @property
    def launch_count(self) -> int:
        return int(self.lib.tim_launch_count(self._ctx))

    @property
    def fold_active(self) -> bool:
        """True while the encoder LayerNorms run folded into the GEMMs (see tim_fold_active in include/tim_b200.h)."""
        return bool(self.lib.tim_fold_active(self._ctx))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.tim_workspace_bytes(self._ctx))


# =======================================================================================================
# drop-in behind an existing reference TIM instance
# =======================================================================================================
def config_from_model(model: torch.nn.Module) -> TIMConfig:
    """Recover the constructor arguments from a reference TIM instance (recognition or detection)."""
    variant = DETECTION if hasattr(model, "backbone") else RECOGNITION
    num_feats = model.feature_encoding.num_feats          # per modality (tim.py:87 doubles model.num_feats only)
    return TIMConfig(num_class=model.num_class, visual_input_dim=model.visual_input_dim,
                     audio_input_dim=model.audio_input_dim, d_model=model.d_model,
                     feedforward_scale=model.dim_feedforward // model.d_model, nhead=model.nhead,
                     num_layers=model.num_layers, input_modality=model.input_modality,
                     data_modality=model.data_modality, num_feats=num_feats,
                     include_verb_noun=bool(model.include_verb_noun), variant=variant)


class _Binding:
    """Keeps the engine's packed weights in sync with the module's parameters (re-packs a parameter when its
    torch version counter moved, e.g. after optimizer.step() or load_state_dict)."""

    def __init__(self, model, engine: TIMEngine):
        self.engine = engine
        self.versions: Dict[str, tuple] = {}
        self.model = model
        self.training_ready = False

    def sync(self):
        sd = {k: v for k, v in self.model.named_parameters()}
        for k in self.engine._keys:
            p = sd[k]
            tag = (p._version, p.data_ptr())
            if self.versions.get(k) != tag:
                self.engine.set_weight(k, p)
                self.versions[k] = tag

    # ---- training leg -------------------------------------------------------------------------------------------
    def enable_training(self):
        """First training-mode call: transposed weight copies in the library (weights are re-packed), ONE flat fp32 gradient buffer
        holding every parameter's .grad (tim_b200.dist.FlatGrads), an autograd anchor, and - when torch.distributed is up - the
        library's NCCL communicator for the single gradient all-reduce that replaces DDP's buckets (models/build.py:58-63)."""
        if self.training_ready:
            return
        import torch.distributed as dist
        from .dist import FlatGrads
        self.engine.enable_training()
        self.versions.clear()                       # every weight is packed again, now with its transposed copy
        self.flat = FlatGrads(self.model.named_parameters())
        dev = self.engine.device
        self.anchor = torch.zeros(1, device=dev, requires_grad=True)
        self.bound = {}
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.distributed:
            self.engine.comm_init()
        self.callback_queued = False
        self.drop_step = 0
        self.training_ready = True

    def dropout_probs(self):
        """(feat_drop, seq_drop, enc_dropout) as the reference's modules hold them NOW (helpers/encodings.py:141,149,177,
        helpers/transformers.py:73-82) - read at every step, so a schedule that edits module.p keeps working. The library has one
        probability for the transformer's four sites in all layers, which is how the reference constructs them (tim.py:119)."""
        fe = self.model.feature_encoding
        feat, enc = set(), set()
        for name in ("visual_embedder", "audio_embedder"):
            first = getattr(getattr(fe, name, None), "0", None)             # nn.Sequential(Dropout, Linear, GELU, LayerNorm)
            if isinstance(first, torch.nn.Dropout):
                feat.add(float(first.p))
        for layer in getattr(self.model, self.engine.cfg.encoder_prefix).layers.children():
            enc.add(float(getattr(layer.self_attn, "dropout", 0.0)))
            enc.update(float(m.p) for m in (getattr(layer, n, None) for n in ("dropout1", "dropout", "dropout2")) if m is not None)
        if len(feat) > 1 or len(enc) > 1:
            raise NotImplementedError(f"tim_b200: the dropout probabilities differ between sites (embedders {sorted(feat)}, encoder "
                                      f"layers {sorted(enc)}); the library takes one feat_drop and one enc_dropout")
        seq = getattr(fe, "dropout", None)
        return (feat.pop() if feat else 0.0), (float(seq.p) if seq is not None else 0.0), (enc.pop() if enc else 0.0)

    def apply_dropout(self):
        """Dropout of the next training forward: the module's probabilities and a fresh seed. The seed is a function of
        torch.initial_seed() (what torch.manual_seed set), the rank and a step counter - reproducible, different on every rank and
        step, and it consumes nothing from torch's generators (the reference's CPU randperm draws stay where they were)."""
        p_feat, p_seq, p_enc = self.dropout_probs() if self.model.training else (0.0, 0.0, 0.0)
        import torch.distributed as dist
        rank = dist.get_rank() if self.distributed else 0
        x = (torch.initial_seed() + 0x9E3779B97F4A7C15 * (self.drop_step + 1) + (rank << 40)) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF          # splitmix64 finaliser
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 31
        self.drop_step += 1
        self.engine.set_dropout(p_feat, p_seq, p_enc, x)

    def attach_grads(self):
        """param.grad must be the flat buffer's views when the backward runs (optimizer.zero_grad() sets them to None by default):
        re-attach (a detached gradient counts as zero) and (re-)bind the library's destinations."""
        params = dict(self.model.named_parameters())
        self.flat.reattach()
        for k in self.engine._keys:
            v = self.flat.view(k).view_as(params[k])
            if self.bound.get(k) != v.data_ptr():
                self.engine.bind_grad(k, v)
                self.bound[k] = v.data_ptr()

    def queue_allreduce(self):
        """Runs once, after the WHOLE backward pass (the callback mechanism DDP's reducer uses): the single all-reduce over the flat
        gradient buffer, enqueued on the current stream behind the last gradient kernel."""
        if not self.distributed or self.callback_queued:
            return
        self.callback_queued = True

        def _cb():
            self.callback_queued = False
            self.engine.allreduce(self.flat.buffer)
        torch.autograd.Variable._execution_engine.queue_callback(_cb)


class _TimeMLPFn(torch.autograd.Function):
    """forward_type == "time_mlp" in training mode: tim_time_mlp_fwd_train / tim_time_mlp_bwd."""

    @staticmethod
    def forward(ctx, binding, times, anchor):
        ctx.binding = binding
        return binding.engine.time_mlp_train(times)

    @staticmethod
    def backward(ctx, d_out):
        b = ctx.binding
        b.queue_allreduce()
        b.engine.time_mlp_bwd(d_out)
        return None, None, None


class _EncoderFn(torch.autograd.Function):
    """forward_type == "encoder" in training mode: tim_encoder_fwd_train / tim_encoder_bwd. Parameter gradients do not travel through
    autograd: the library accumulates them straight into param.grad (views of the flat buffer)."""
    KEYS = ("verb", "noun", "action", "audio", "reg_v", "reg_a", "feats")

    @staticmethod
    def forward(ctx, binding, vis, aud, time_enc, Qv, Qa, anchor):
        o = binding.engine.encoder_train(vis, aud, time_enc, Qv, Qa)
        ctx.binding = binding
        return tuple(o[k] for k in _EncoderFn.KEYS)          # None where the reference returns None (non-tensor outputs pass through)

    @staticmethod
    def backward(ctx, *grads):
        b = ctx.binding
        b.queue_allreduce()
        d_te = b.engine.encoder_bwd(dict(zip(_EncoderFn.KEYS, grads)))
        return None, None, None, d_te, None, None, None


def _encoder_train(b: "_Binding", vis, aud, time_enc, Qv, Qa):
    b.enable_training()
    b.sync()
    b.attach_grads()
    b.apply_dropout()
    outs = _EncoderFn.apply(b, vis, aud, time_enc, int(Qv or 0), int(Qa or 0), b.anchor)
    return dict(zip(_EncoderFn.KEYS, outs))


def _forward_recognition(self, inputs, forward_type, time_encodings=None, num_v_queries=None, num_a_queries=None):
    """Same signature / returns as recognition/.../models/tim.py:174-191."""
    b: _Binding = self._tim_b200
    if forward_type == "drloc_mlp":                        # loss-side helper, stays in PyTorch
        return self.drloc_mlp(inputs).squeeze(2)
    if getattr(self, "pool_features", False):
        raise NotImplementedError("tim_b200: AVGA feature pooling (AVE-only) is outside the hot path")
    if forward_type not in ("time_mlp", "encoder"):
        raise ValueError(f"unknown forward_type {forward_type!r}")
    if self.training:
        # the reference applies dropout whenever the module is in train() mode, with or without autograd (helpers/transformers.py:
        # 73-82, encodings.py:141,149,177): the encoder of a train()-mode module always runs the training forward (tim_set_dropout)
        b.enable_training()
        if forward_type == "time_mlp" and torch.is_grad_enabled():
            b.sync()
            b.attach_grads()
            return _TimeMLPFn.apply(b, inputs, b.anchor)
        if forward_type == "encoder":
            o = _encoder_train(b, inputs[0], inputs[1], time_encodings, num_v_queries, num_a_queries)
            return (o["verb"], o["noun"], o["action"], o["audio"]), o["feats"]
    b.sync()
    if forward_type == "time_mlp":
        return b.engine.time_mlp(inputs)
    if forward_type == "encoder":
        o = b.engine.encoder(inputs[0], inputs[1], time_encodings, num_v_queries, num_a_queries)
        if b.engine.fold_check():           # the folded-LayerNorm precision guard fired: redo with the un-folded flow, now
            o = b.engine.encoder(inputs[0], inputs[1], time_encodings, num_v_queries, num_a_queries)
        return (o["verb"], o["noun"], o["action"], o["audio"]), o["feats"]
    raise ValueError(f"unknown forward_type {forward_type!r}")


def _label_queries_device(model, engine: "TIMEngine", queries, target, modality):
    """label_queries + assign_positive_labels of the reference (detection/.../models/tim.py:158-270) through the library: same
    inputs (the target dict of the data loader), same return structure."""
    dev = queries.device
    if modality == "visual":
        segs = target["v_gt_segments"]
        labels = torch.stack([target["verb"], target["noun"], target["action"]], dim=-1)
    else:
        segs = target["a_gt_segments"]
        labels = target["class_id"].unsqueeze(-1)
    targets, ids, ious = engine.label_queries(queries, segs, labels, model.iou_threshold)
    s = model.label_smoothing
    if modality == "visual":
        verb = noun = torch.empty(size=(0,), device=dev)
        n_act = model.num_class[0]
        if model.include_verb_noun:
            n_verb, n_noun, n_act = model.num_class[0][0], model.num_class[0][1], model.num_class[0][2]
            verb, noun = engine.smooth_labels(ids, 0, n_verb, s), engine.smooth_labels(ids, 1, n_noun, s)
        query_labels = [verb, noun, engine.smooth_labels(ids, 2, n_act, s)]
    else:
        query_labels = engine.smooth_labels(ids, ids.shape[1] - 1, model.num_class[1], s)
    return targets, query_labels, ious


def _forward_detection(self, inputs, forward_type, feature_times=None, target=None, label_queries=False):
    """Same signature / returns as detection/.../models/tim.py:415-430 (inference branch :339-400)."""
    b: _Binding = self._tim_b200
    if forward_type == "drloc_mlp":
        return self.drloc_mlp(inputs).squeeze(2)
    if forward_type != "encoder":
        raise ValueError(f"unknown forward_type {forward_type!r}")
    if self.training:
        return _forward_detection_train(self, inputs, feature_times, target)
    b.sync()
    cfg = b.engine.cfg
    dev = feature_times.device
    v_offsets = a_offsets = torch.empty(0, 2)
    v_labels = a_labels = torch.empty(0, 4)
    nv = na = 0
    v_queries = a_queries = v_ious = a_ious = None
    all_times = feature_times
    B = all_times.shape[0]
    if "visual" in cfg.data_modality:
        v_queries = self.inference_queries.repeat(B, 1, 1).to(device=dev)
        nv = v_queries.shape[1]
        if label_queries:                                  # IoU / arg-max / smoothed labels on the device (labels.cu)
            v_offsets, v_labels, v_ious = _label_queries_device(self, b.engine, v_queries, target, "visual")
        all_times = torch.cat([all_times, v_queries], dim=1)
        v_queries = torch.flatten(v_queries, 0, 1)
    if "audio" in cfg.data_modality:
        # the reference draws a permutation here and discards it (detection/.../tim.py:364): one draw from torch's default CPU generator per
        # audio-modality inference forward. Reproduced so that whatever the caller draws next sees the generator state it would see there.
        torch.randperm(self.inference_queries.shape[1])
        a_queries = self.inference_queries.repeat(B, 1, 1).to(device=dev)
        na = a_queries.shape[1]
        if label_queries:
            a_offsets, a_labels, a_ious = _label_queries_device(self, b.engine, a_queries, target, "audio")
        all_times = torch.cat([all_times, a_queries], dim=1)
        a_queries = torch.flatten(a_queries, 0, 1)
    te = b.engine.time_mlp(all_times)
    o = b.engine.encoder(inputs[0], inputs[1], te, nv, na)
    if b.engine.fold_check():
        o = b.engine.encoder(inputs[0], inputs[1], te, nv, na)
    cls = (o["verb"], o["noun"], o["action"], o["audio"])
    reg = (o["reg_v"], o["reg_a"])
    return (cls, reg, o["feats"]), (v_offsets, a_offsets), (v_labels, a_labels), (v_queries, a_queries), (v_ious, a_ious)


def _forward_detection_train(self, inputs, feature_times, target):
    """detection/.../models/tim.py:272-337 (forward_train): queries drawn from the train pool with the CPU global RNG exactly as the
    reference does (torch.randperm, one draw per modality), labelled on the device, then time_mlp + encoder through the library."""
    b: _Binding = self._tim_b200
    b.enable_training()
    cfg = b.engine.cfg
    dev = feature_times.device
    v_offsets = a_offsets = torch.empty(0, 2)
    v_labels = a_labels = torch.empty(0, 4)
    nv = na = 0
    v_queries = a_queries = v_ious = a_ious = None
    all_times = feature_times
    B = all_times.shape[0]
    if "visual" in cfg.data_modality:
        idx = torch.randperm(self.train_pool.shape[1])[:self.num_queries]
        v_queries = self.train_pool[:, idx.long()].repeat(B, 1, 1).to(device=dev)
        nv = v_queries.shape[1]
        v_offsets, v_labels, v_ious = _label_queries_device(self, b.engine, v_queries, target, "visual")
        all_times = torch.cat([all_times, v_queries], dim=1)
        v_queries = torch.flatten(v_queries, 0, 1)
    if "audio" in cfg.data_modality:
        idx = torch.randperm(self.train_pool.shape[1])[:self.num_queries]
        a_queries = self.train_pool[:, idx.long()].repeat(B, 1, 1).to(device=dev)
        na = a_queries.shape[1]
        a_offsets, a_labels, a_ious = _label_queries_device(self, b.engine, a_queries, target, "audio")
        all_times = torch.cat([all_times, a_queries], dim=1)
        a_queries = torch.flatten(a_queries, 0, 1)
    b.sync()
    if torch.is_grad_enabled():
        b.attach_grads()
        te = _TimeMLPFn.apply(b, all_times, b.anchor)
    else:
        te = b.engine.time_mlp(all_times)
    o = _encoder_train(b, inputs[0], inputs[1], te, nv, na)       # train() mode: dropout applies with or without autograd
    cls = (o["verb"], o["noun"], o["action"], o["audio"])
    reg = (o["reg_v"], o["reg_a"])
    return (cls, reg, o["feats"]), (v_offsets, a_offsets), (v_labels, a_labels), (v_queries, a_queries), (v_ious, a_ious)


def patch_model(model: torch.nn.Module, compute_dtype: str = "fp16", device: Optional[int] = None) -> torch.nn.Module:
    """Route model.forward through libtim_b200. `model` is a reference TIM (or DDP(module=TIM)) already on a GPU."""
    inner = model.module if hasattr(model, "module") and not hasattr(model, "time_mlp") else model
    p = next(inner.parameters())
    if p.device.type != "cuda":
        raise RuntimeError("tim_b200.patch_model: the model must be on a CUDA device (no CPU fallback)")
    dev = p.device.index if device is None else device
    cfg = config_from_model(inner)
    engine = TIMEngine(cfg, dev, compute_dtype)
    inner._tim_b200 = _Binding(inner, engine)
    fwd = _forward_recognition if cfg.variant == RECOGNITION else _forward_detection
    inner.forward = types.MethodType(fwd, inner)
    return model
