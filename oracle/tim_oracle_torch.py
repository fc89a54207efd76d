"""The reference forward restated with the SAME PyTorch CPU calls the reference's modules make - the TIMED CPU arm of bench.py
(`--impl reference` and the `cpu_baseline` object). TEST / MEASUREMENT INFRASTRUCTURE ONLY: the product path (tim_b200/) never
imports it; parity is judged by the torch-independent numpy oracle (oracle/tim_oracle.py), against which this file is checked.

Why a second restatement: the reference's arithmetic lives in PyTorch (nn.Linear / nn.LayerNorm / nn.GELU / nn.MultiheadAttention,
SURVEY.md §8c) and its Python package cannot travel to the GPU box. A numpy port computes the same numbers but is 2.9x SLOWER than
the reference itself on the same cores (measured in the build container on cfg2, 24 clips, 8 threads: reference 973 ms, numpy port
2788 ms, this file 980 ms), which would flatter the GPU / CPU ratio. This file issues what the reference issues, op for op:

  time MLP                 recognition/time_interval_machine/models/tim.py:66-74          F.linear + relu x3, F.layer_norm
  embedders, token concat  .../models/helpers/encodings.py:140-153,181-251               F.linear, F.gelu, F.layer_norm, torch.cat,
                           (detection: .../helpers/encodings.py:152-201)                  transpose(0, 1).contiguous() to [S, B, E]
  mask                     .../models/tim.py:161-166                                      the [B*H, S, S] boolean mask IS materialised
                                                                                          with repeat_interleave, as the reference does
  encoder layer            .../models/helpers/transformers.py:92-111                      F.multi_head_attention_forward (the slow path
                                                                                          nn.MultiheadAttention takes with batch_first
                                                                                          False and need_weights=True), F.layer_norm,
                                                                                          F.linear, F.gelu
  output transpose         .../models/helpers/transformers.py:32-48                       transpose(0, 1).contiguous()
  heads                    .../models/helpers/head.py (recognition / detection)           F.linear on the row slices, sigmoid MLP

Checked by tests/test_oracle_golden.py against the golden vectors minted from the real reference (<= 2e-6 rel-L2).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from tim_b200.config import RECOGNITION, TIMConfig


class TIMOracleTorch:
    def __init__(self, cfg: TIMConfig, sd: Dict[str, np.ndarray], dtype=torch.float32, device="cpu"):
        """device="cpu" is the timed CPU arm; a CUDA device gives the eager-PyTorch GPU baseline of tools/eager_baseline.py (the same
        library calls the reference would make on a GPU)."""
        self.cfg = cfg
        self.dt = dtype
        self.dev = torch.device(device)
        self.sd = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device=self.dev, dtype=dtype) for k, v in sd.items()}

    def time_mlp(self, times: torch.Tensor) -> torch.Tensor:
        s = self.sd
        x = times
        for i in (0, 2, 4):
            x = F.relu(F.linear(x, s[f"time_mlp.{i}.weight"], s[f"time_mlp.{i}.bias"]))
        return F.layer_norm(x, (x.shape[-1],), s["time_mlp.6.weight"], s["time_mlp.6.bias"], 1e-5)

    def _embed(self, x, which):
        s, p = self.sd, f"feature_encoding.{which}_embedder."
        x = F.gelu(F.linear(x, s[p + "1.weight"], s[p + "1.bias"]))
        return F.layer_norm(x, (x.shape[-1],), s[p + "3.weight"], s[p + "3.bias"], 1e-5)

    def assemble(self, vis, aud, te, Qv, Qa) -> torch.Tensor:
        """[S, B, E], as the reference's feature_encoding returns it."""
        cfg, s = self.cfg, self.sd
        Fn, fe, B = cfg.num_feats, "feature_encoding.", te.shape[0]

        def cls_block(param, te_slice, n):
            return torch.cat([s[fe + param].expand(B, n, -1), te_slice], dim=-1)

        if cfg.input_modality == "audio_visual":
            vmod, amod = s[fe + "visual_modality_encoding"], s[fe + "audio_modality_encoding"]
            seq = [torch.cat([self._embed(vis, "visual"), te[:, :Fn]], dim=-1) + vmod,
                   torch.cat([self._embed(aud, "audio"), te[:, Fn:2 * Fn]], dim=-1) + amod]
            qte = te[:, 2 * Fn:]
            if "visual" in cfg.data_modality and Qv > 0:
                if cfg.verb_noun_tokens:
                    seq.append(cls_block("visual_verb_cls", qte[:, :Qv], Qv) + vmod)
                    seq.append(cls_block("visual_noun_cls", qte[:, :Qv], Qv) + vmod)
                seq.append(cls_block("visual_action_cls", qte[:, :Qv], Qv) + vmod)
            if "audio" in cfg.data_modality and Qa > 0:
                seq.append(cls_block("audio_action_cls", qte[:, qte.shape[1] - Qa:], Qa) + amod)
        elif cfg.input_modality == "visual":
            seq = [torch.cat([self._embed(vis, "visual"), te[:, :Fn]], dim=-1)]
            qte = te[:, Fn:]
            if cfg.variant == RECOGNITION:
                if cfg.include_verb_noun:
                    seq += [cls_block("verb_cls", qte, Qv), cls_block("noun_cls", qte, Qv)]
                seq.append(cls_block("action_cls", qte, Qv))
            else:
                seq.append(cls_block("visual_action_cls", qte, Qv))
        else:
            seq = [torch.cat([self._embed(aud, "audio"), te[:, :Fn]], dim=-1),
                   cls_block("action_cls" if cfg.variant == RECOGNITION else "audio_action_cls", te[:, Fn:], Qa)]
        return torch.cat(seq, dim=1).transpose(0, 1).contiguous()

    def backbone(self, x: torch.Tensor) -> torch.Tensor:
        """x [S, B, E] -> [B, S, E]; the reference's mask construction and encoder loop (tim.py:161-168, transformers.py:32-48)."""
        cfg, s = self.cfg, self.sd
        S, B, E = x.shape
        masks = torch.ones((S, S), device=x.device)
        masks[:, :cfg.F_tot] = 0.
        masks = masks.fill_diagonal_(0.).unsqueeze(0)
        masks = masks.repeat_interleave(cfg.nhead * B, dim=0).bool()
        for l in range(cfg.num_layers):
            p = f"{cfg.encoder_prefix}.layers.{l}."
            a, _ = F.multi_head_attention_forward(
                x, x, x, E, cfg.nhead, s[p + "self_attn.in_proj_weight"], s[p + "self_attn.in_proj_bias"], None, None, False, 0.0,
                s[p + "self_attn.out_proj.weight"], s[p + "self_attn.out_proj.bias"], training=False, key_padding_mask=None,
                need_weights=True, attn_mask=masks)
            x = F.layer_norm(x + a, (E,), s[p + "norm1.weight"], s[p + "norm1.bias"], 1e-5)
            h = F.linear(F.gelu(F.linear(x, s[p + "linear1.weight"], s[p + "linear1.bias"])), s[p + "linear2.weight"], s[p + "linear2.bias"])
            x = F.layer_norm(x + h, (E,), s[p + "norm2.weight"], s[p + "norm2.bias"], 1e-5)
        return x.transpose(0, 1).contiguous()

    def heads(self, x: torch.Tensor, Qv: int, Qa: int) -> Dict[str, Optional[torch.Tensor]]:
        cfg, s = self.cfg, self.sd
        hc, S = cfg.head_classes(), x.shape[1]
        out: Dict[str, Optional[torch.Tensor]] = {k: None for k in ("verb", "noun", "action", "audio", "reg_v", "reg_a")}

        def fc(name, rows):
            return F.linear(rows, s[f"cls_head.{name}.weight"], s[f"cls_head.{name}.bias"]).flatten(0, 1)

        def reg(name, rows):
            p = f"reg_head.{name}."
            y = F.relu(F.linear(rows, s[p + "0.weight"], s[p + "0.bias"]))
            y = F.relu(F.linear(y, s[p + "2.weight"], s[p + "2.bias"]))
            return torch.sigmoid(F.linear(y, s[p + "4.weight"], s[p + "4.bias"])).flatten(0, 1)

        has_v, has_a = "visual" in cfg.data_modality, "audio" in cfg.data_modality
        aud_start = S - Qa if (has_a and Qa > 0) else S
        if cfg.variant == RECOGNITION:
            act_start = aud_start - Qv
            if has_v:
                if hc["verb"]:
                    out["verb"] = fc("fc_visual_verb", x[:, act_start - 2 * Qv:act_start - Qv])
                    out["noun"] = fc("fc_visual_noun", x[:, act_start - Qv:act_start])
                out["action"] = fc("fc_visual_action", x[:, act_start:aud_start])
            if has_a:
                out["audio"] = fc("fc_audio_action", x[:, aud_start:])
        else:
            vis_start = aud_start - Qv
            if has_v:
                rows = x[:, vis_start:aud_start]
                if hc["verb"]:
                    out["verb"], out["noun"] = fc("fc_visual_verb", rows), fc("fc_visual_noun", rows)
                out["action"], out["reg_v"] = fc("fc_visual_action", rows), reg("fc_visual_action", rows)
            if has_a:
                rows = x[:, aud_start:]
                out["audio"], out["reg_a"] = fc("fc_audio_action", rows), reg("fc_audio_action", rows)
        return out

    @torch.no_grad()
    def forward(self, vis, aud, times, Qv: int, Qa: int) -> Dict[str, Optional[np.ndarray]]:
        """numpy in, numpy out (dict as TIMOracle.forward): time_mlp + encoder on raw interval times [B, T, 2]."""
        t = (lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(device=self.dev, dtype=self.dt))
        out = self.forward_tensors(t(vis), t(aud), t(times), Qv, Qa)
        return {k: (None if v is None else v.float().cpu().numpy()) for k, v in out.items()}

    @torch.no_grad()
    def forward_tensors(self, vis, aud, times, Qv: int, Qa: int) -> Dict[str, Optional[torch.Tensor]]:
        """Tensors (already on self.dev) in, tensors out - no host round trip (what tools/eager_baseline.py times)."""
        te = self.time_mlp(times)
        x = self.backbone(self.assemble(vis, aud, te, Qv, Qa))
        out = self.heads(x, Qv, Qa)
        out["feats"] = x[:, :self.cfg.F_tot]
        out["time_encodings"] = te
        return out
