"""CPU restatement (numpy) of the reference TIM forward — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product path (tim_b200/) never does; it fails loudly if the CUDA
library is missing.

What it restates (dense S x S masked attention exactly as the reference executes it):
  time MLP                     recognition/time_interval_machine/models/tim.py:66-74
  feature / CLS token assembly recognition/.../models/helpers/encodings.py:41-75,102-121,181-251
                               detection/.../models/helpers/encodings.py:33-53,81-100,152-201
  mask                         recognition/.../models/tim.py:161-166 ; detection/.../tim.py:384-389
  encoder layer (post-LN)      */models/helpers/transformers.py:92-111
  MHA arithmetic               torch.nn.functional.multi_head_attention_forward (pinned pytorch 1.11.0,
                               environment.yml:13; third-party, not vendored): packed in-proj [q;k;v],
                               q scaled by head_dim**-0.5 before q.k^T, boolean mask True -> -inf,
                               softmax over keys, exact-erf GELU, LayerNorm biased variance eps=1e-5
  CLS heads                    recognition/.../helpers/head.py:17-38,57-69,76-81
                               detection/.../helpers/head.py:27-46,65-79,89-93
  regression heads             detection/.../helpers/head.py:116-126,141-145,159-163
  forward_encoder glue         recognition/.../tim.py:147-172 ; detection/.../tim.py:339-400

Parity pin: the reference ships NO golden vectors or numeric tests, so this oracle is
pinned against outputs of the reference itself, imported from /root/reference and run on CPU fp32 by
tools/make_golden.py; those outputs are committed under tests/golden/ and checked by
tests/test_oracle_golden.py.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

try:  # exact erf; scipy is in the image. math.erf fallback keeps the oracle dependency-light.
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    import math
    _erf = np.vectorize(math.erf, otypes=[np.float64])

from tim_b200.config import TIMConfig, RECOGNITION


def _linear(x, w, b):
    return x @ w.T + b


def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True)
    return xc / np.sqrt(var + x.dtype.type(eps)) * w + b


def _gelu(x):
    return (0.5 * x * (1.0 + _erf(x / np.sqrt(2.0).astype(x.dtype)))).astype(x.dtype)


def _relu(x):
    return np.maximum(x, 0)


class TIMOracle:
    """Functional numpy TIM. `sd` maps reference state_dict keys to arrays."""

    def __init__(self, cfg: TIMConfig, sd: Dict[str, np.ndarray], dtype=np.float32):
        self.cfg = cfg
        self.dt = np.dtype(dtype)
        self.sd = {k: np.asarray(v, dtype=self.dt) for k, v in sd.items()}

    # ---- tim.py:66-74 ----
    def time_mlp(self, times: np.ndarray) -> np.ndarray:
        s = self.sd
        x = np.asarray(times, self.dt)
        x = _relu(_linear(x, s["time_mlp.0.weight"], s["time_mlp.0.bias"]))
        x = _relu(_linear(x, s["time_mlp.2.weight"], s["time_mlp.2.bias"]))
        x = _relu(_linear(x, s["time_mlp.4.weight"], s["time_mlp.4.bias"]))
        return _layer_norm(x, s["time_mlp.6.weight"], s["time_mlp.6.bias"])

    def _embed(self, x, which):
        s = self.sd
        p = f"feature_encoding.{which}_embedder."
        x = _gelu(_linear(np.asarray(x, self.dt), s[p + "1.weight"], s[p + "1.bias"]))
        return _layer_norm(x, s[p + "3.weight"], s[p + "3.bias"])

    # ---- encodings.py forward() of the three encoding classes; returns [B, S, E] ----
    def assemble(self, vis, aud, te, Qv, Qa) -> np.ndarray:
        cfg, s = self.cfg, self.sd
        F = cfg.num_feats
        fe = "feature_encoding."
        te = np.asarray(te, self.dt)
        B = te.shape[0]

        def cls_block(param, te_slice, n):
            c = np.broadcast_to(s[fe + param], (B, n, cfg.d_model))
            return np.concatenate([c, te_slice], axis=-1)

        if cfg.input_modality == "audio_visual":
            vmod, amod = s[fe + "visual_modality_encoding"], s[fe + "audio_modality_encoding"]
            v = np.concatenate([self._embed(vis, "visual"), te[:, :F]], axis=-1) + vmod
            a = np.concatenate([self._embed(aud, "audio"), te[:, F:2 * F]], axis=-1) + amod
            seq = [v, a]
            qte = te[:, 2 * F:]
            if "visual" in cfg.data_modality and Qv > 0:
                if cfg.verb_noun_tokens:
                    seq.append(cls_block("visual_verb_cls", qte[:, :Qv], Qv) + vmod)
                    seq.append(cls_block("visual_noun_cls", qte[:, :Qv], Qv) + vmod)
                seq.append(cls_block("visual_action_cls", qte[:, :Qv], Qv) + vmod)
            if "audio" in cfg.data_modality and Qa > 0:
                seq.append(cls_block("audio_action_cls", qte[:, qte.shape[1] - Qa:], Qa) + amod)
            return np.concatenate(seq, axis=1)
        if cfg.input_modality == "visual":
            seq = [np.concatenate([self._embed(vis, "visual"), te[:, :F]], axis=-1)]
            qte = te[:, F:]
            if cfg.variant == RECOGNITION:
                if cfg.include_verb_noun:
                    seq.append(cls_block("verb_cls", qte, Qv))
                    seq.append(cls_block("noun_cls", qte, Qv))
                seq.append(cls_block("action_cls", qte, Qv))
            else:
                seq.append(cls_block("visual_action_cls", qte, Qv))
            return np.concatenate(seq, axis=1)
        seq = [np.concatenate([self._embed(aud, "audio"), te[:, :F]], axis=-1)]
        name = "action_cls" if cfg.variant == RECOGNITION else "audio_action_cls"
        seq.append(cls_block(name, te[:, F:], Qa))
        return np.concatenate(seq, axis=1)

    # ---- transformers.py:92-111 with the tim.py:161-166 mask; x is [B, S, E] ----
    def encoder_layer(self, x, l: int, mask: np.ndarray) -> np.ndarray:
        cfg, s = self.cfg, self.sd
        p = f"{cfg.encoder_prefix}.layers.{l}."
        B, S, E = x.shape
        H, hd = cfg.nhead, cfg.head_dim
        qkv = _linear(x, s[p + "self_attn.in_proj_weight"], s[p + "self_attn.in_proj_bias"])
        q, k, v = qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:]
        q = q * self.dt.type(hd ** -0.5)
        q = q.reshape(B, S, H, hd).transpose(0, 2, 1, 3)
        k = k.reshape(B, S, H, hd).transpose(0, 2, 1, 3)
        v = v.reshape(B, S, H, hd).transpose(0, 2, 1, 3)
        sc = q @ k.transpose(0, 1, 3, 2)                      # [B,H,S,S] dense, as the reference does
        sc = np.where(mask[None, None], self.dt.type(-np.inf), sc)
        sc = sc - sc.max(axis=-1, keepdims=True)
        pr = np.exp(sc)
        pr = pr / pr.sum(axis=-1, keepdims=True)
        a = (pr @ v).transpose(0, 2, 1, 3).reshape(B, S, E)
        a = _linear(a, s[p + "self_attn.out_proj.weight"], s[p + "self_attn.out_proj.bias"])
        x = _layer_norm(x + a, s[p + "norm1.weight"], s[p + "norm1.bias"])
        h = _gelu(_linear(x, s[p + "linear1.weight"], s[p + "linear1.bias"]))
        h = _linear(h, s[p + "linear2.weight"], s[p + "linear2.bias"])
        return _layer_norm(x + h, s[p + "norm2.weight"], s[p + "norm2.bias"])

    def mask(self, S: int) -> np.ndarray:
        """True = masked. mask[i,j] = (j >= num_feats) and (i != j)   (tim.py:161-166)."""
        m = np.ones((S, S), dtype=bool)
        m[:, :self.cfg.F_tot] = False
        np.fill_diagonal(m, False)
        return m

    def backbone(self, x: np.ndarray, clip_chunk: int = 8) -> np.ndarray:
        m = self.mask(x.shape[1])
        outs = []
        for b0 in range(0, x.shape[0], clip_chunk):         # chunk only to bound the S x S memory
            xb = x[b0:b0 + clip_chunk]
            for l in range(self.cfg.num_layers):
                xb = self.encoder_layer(xb, l, m)
            outs.append(xb)
        return np.concatenate(outs, axis=0)

    # ---- head.py ----
    def heads(self, x: np.ndarray, Qv: int, Qa: int) -> Dict[str, Optional[np.ndarray]]:
        cfg, s = self.cfg, self.sd
        hc = cfg.head_classes()
        S = x.shape[1]
        out: Dict[str, Optional[np.ndarray]] = {k: None for k in ("verb", "noun", "action", "audio", "reg_v", "reg_a")}

        def fc(name, rows):
            w, b = s[f"cls_head.{name}.weight"], s[f"cls_head.{name}.bias"]
            y = _linear(rows, w, b)
            return y.reshape(-1, y.shape[-1])

        def reg(name, rows):
            p = f"reg_head.{name}."
            y = _relu(_linear(rows, s[p + "0.weight"], s[p + "0.bias"]))
            y = _relu(_linear(y, s[p + "2.weight"], s[p + "2.bias"]))
            y = _linear(y, s[p + "4.weight"], s[p + "4.bias"])
            y = 1.0 / (1.0 + np.exp(-y))
            return y.reshape(-1, 2).astype(self.dt)

        has_v = "visual" in cfg.data_modality
        has_a = "audio" in cfg.data_modality
        aud_start = S - Qa if (has_a and Qa > 0) else S
        if cfg.variant == RECOGNITION:
            act_start = aud_start - Qv
            if has_v:
                if hc["verb"]:
                    noun_start = act_start - Qv
                    verb_start = noun_start - Qv
                    out["verb"] = fc("fc_visual_verb", x[:, verb_start:noun_start])
                    out["noun"] = fc("fc_visual_noun", x[:, noun_start:act_start])
                out["action"] = fc("fc_visual_action", x[:, act_start:aud_start])
            if has_a:
                out["audio"] = fc("fc_audio_action", x[:, aud_start:])
        else:
            vis_start = aud_start - Qv
            if has_v:
                rows = x[:, vis_start:aud_start]
                if hc["verb"]:
                    out["verb"] = fc("fc_visual_verb", rows)
                    out["noun"] = fc("fc_visual_noun", rows)
                out["action"] = fc("fc_visual_action", rows)
                out["reg_v"] = reg("fc_visual_action", rows)
            if has_a:
                rows = x[:, aud_start:]
                out["audio"] = fc("fc_audio_action", rows)
                out["reg_a"] = reg("fc_audio_action", rows)
        return out

    # ---- tim.py forward_encoder ----
    def encoder(self, vis, aud, time_encodings, Qv: int, Qa: int, clip_chunk: int = 8):
        """Returns dict(verb, noun, action, audio, reg_v, reg_a, feats)."""
        x = self.assemble(vis, aud, time_encodings, Qv, Qa)
        x = self.backbone(x, clip_chunk)
        out = self.heads(x, Qv, Qa)
        out["feats"] = x[:, :self.cfg.F_tot]
        return out

    def forward(self, vis, aud, times, Qv: int, Qa: int, clip_chunk: int = 8):
        """time_mlp + encoder on raw interval times [B, T, 2] (what detection does inside one call,
        detection/.../tim.py:378-394, and what the recognition drivers do as two calls)."""
        te = self.time_mlp(times)
        out = self.encoder(vis, aud, te, Qv, Qa, clip_chunk)
        out["time_encodings"] = te
        return out
