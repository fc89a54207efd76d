"""CPU restatement (numpy) of the BACKWARD of the reference TIM forward - TEST INFRASTRUCTURE ONLY (SURVEY.md §8f row 1,
DESIGN.md §9): the checker of the CUDA training leg, never on its path. Same import rules as oracle/tim_oracle.py.

What it differentiates: exactly the graph oracle/tim_oracle.py restates (time MLP -> token assembly -> L post-LN encoder layers with
dense masked attention -> CLS / regression heads, plus the feature rows returned for the drloc loss), i.e. what autograd records
when the reference runs  recognition/.../models/tim.py:147-172  /  detection/.../models/tim.py:339-400. Dropout: the
reference's six nn.Dropout sites (helpers/encodings.py:141,149,177; helpers/transformers.py:73-82,102-108) take their masks from
torch's RNG, which no other implementation can reproduce; with `dropout=` given, this oracle applies at the same six sites the
masks of the library's counter-based hash (drop_mask below restates tim_b200/csrc/kernels.h: DropSite), so gradient parity is
checked mask for mask. The reference has no hand-written backward: gradients come from torch.autograd over
  nn.Linear / ReLU / GELU(erf) / LayerNorm      (tim.py:66-74, helpers/encodings.py:140-153, helpers/transformers.py:75-111, helpers/head.py)
  F.multi_head_attention_forward                (q scaled before q.k^T, boolean mask -> -inf, softmax over keys)
  torch.cat / broadcast of the CLS parameters and modality encodings (helpers/encodings.py:181-251).
The scalar that is differentiated is  L = sum_k <output_k, cotangent_k>  for caller-supplied cotangents of every output
(logits, regression outputs, feature rows), which is how any loss reaches this graph.

Parity pin: tests/golden/grads.npz - fingerprints (norm, sum, 512 sampled entries) of the torch.autograd gradients of the UNMODIFIED reference (float64, CPU) for seeded weights,
inputs and cotangents, minted by tools/make_golden_grads.py; tests/test_oracle_golden.py checks this file against them.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from oracle.tim_oracle import TIMOracle, _erf, _gelu, _relu
from tim_b200.config import RECOGNITION


# ---------------------------------------------------------------------------------------------------------------------
# the library's dropout masks (tim_b200/csrc/kernels.h: DropSite / make_drop_site, ptx.cuh: drop_hash)
# ---------------------------------------------------------------------------------------------------------------------
DROP_FEAT_VIS, DROP_FEAT_AUD, DROP_SEQ, DROP_ATTN, DROP_SUB1, DROP_FFN, DROP_SUB2 = 1, 2, 3, 4, 5, 6, 7
DROP_ATTN_KW = 130


def drop_mask(e, p: float, seed: int, site: int, layer: int = 0):
    """Scaled keep mask (float64: 0 or 65536 / (65536 - thr)) of the elements with flat indices e (any shape, < 2^32)."""
    e = np.asarray(e, np.uint64)
    if p <= 0.0:
        return np.ones(e.shape)
    thr = min(int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)), 65535)
    scale = float(np.float32(65536.0) / np.float32(65536 - thr))
    seed32 = (int(seed) ^ (int(seed) >> 32)) & 0xFFFFFFFF
    key = (seed32 ^ (site * 0x9E3779B9) ^ (layer * 0x85EBCA6B)) & 0xFFFFFFFF
    m32 = np.uint64(0xFFFFFFFF)
    x = ((e >> np.uint64(1)) ^ np.uint64(key)) & m32
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & m32
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & m32
    x ^= x >> np.uint64(16)
    bits = np.where((e & np.uint64(1)) != 0, x >> np.uint64(16), x & np.uint64(0xFFFF))
    return np.where(bits >= np.uint64(thr), scale, 0.0)


def two_stream_rows(B: int, S: int, Ft: int):
    """[B, S] -> row of token (b, s) in the library's layout: the B*Ft feature rows first, then the B*(S-Ft) query rows."""
    b = np.arange(B)[:, None]
    sidx = np.arange(S)[None, :]
    Qt = S - Ft
    return np.where(sidx < Ft, b * Ft + sidx, B * Ft + b * Qt + (sidx - Ft))


def row_drop_mask(B: int, S: int, Ft: int, width: int, p: float, seed: int, site: int, layer: int = 0):
    """[B, S, width] mask of a token-row site (seq_drop, dropout1, the FFN's dropout, dropout2): element = library row * width + column."""
    rows = two_stream_rows(B, S, Ft)
    return drop_mask(rows[..., None] * width + np.arange(width), p, seed, site, layer)


def attn_drop_mask(B: int, H: int, S: int, Ft: int, p: float, seed: int, layer: int):
    """[B, H, S, S] mask of the attention probabilities: element = ((b H + h) S + row) * DROP_ATTN_KW + key index, where the key index
    of feature key j is j and a query row's own key is Ft (the other entries are removed by the attention mask anyway)."""
    kidx = np.broadcast_to(np.arange(S)[None, :], (S, S)).copy()
    qrows = np.arange(S) >= Ft
    kidx[qrows, np.arange(S)[qrows]] = Ft
    kidx = np.minimum(kidx, DROP_ATTN_KW - 1)
    base = ((np.arange(B)[:, None, None] * H + np.arange(H)[None, :, None]) * S + np.arange(S)[None, None, :]) * DROP_ATTN_KW
    return drop_mask(base[..., None] + kidx[None, None], p, seed, DROP_ATTN, layer)


def _ln_fwd(x, w, b, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    xc = x - mu
    rstd = 1.0 / np.sqrt((xc * xc).mean(axis=-1, keepdims=True) + x.dtype.type(eps))
    xhat = xc * rstd
    return xhat * w + b, (xhat, rstd)


def _ln_bwd(dy, cache, w):
    xhat, rstd = cache
    dxhat = dy * w
    dx = rstd * (dxhat - dxhat.mean(axis=-1, keepdims=True) - xhat * (dxhat * xhat).mean(axis=-1, keepdims=True))
    red = tuple(range(dy.ndim - 1))
    return dx, (dy * xhat).sum(axis=red), dy.sum(axis=red)


def _gelu_bwd(dy, x):
    dt = x.dtype.type
    cdf = 0.5 * (1.0 + _erf(x / np.sqrt(dt(2.0))))
    pdf = np.exp(-0.5 * x * x) / np.sqrt(dt(2.0 * np.pi))
    return (dy * (cdf + x * pdf)).astype(x.dtype)


def _lin_bwd(dy, x, w):
    """y = x w^T + b -> (dx, dw, db)."""
    dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
    return dy @ w, dy2.T @ x2, dy2.sum(axis=0)


class TIMOracleGrad(TIMOracle):
    """forward_backward(...) -> (outputs, grads): grads maps every reference state_dict key that receives a gradient, plus
    'input.vis' / 'input.aud', to d L / d tensor for L = sum_k <outputs[k], cot[k]>."""

    def forward_backward(self, vis, aud, times, Qv: int, Qa: int, cot: Dict[str, np.ndarray], dropout: Optional[dict] = None):
        """dropout: None, or {"p_feat", "p_seq", "p_enc", "seed"} - the arguments of tim_set_dropout."""
        cfg, s, dt = self.cfg, self.sd, self.dt
        g: Dict[str, np.ndarray] = {}
        dr = dropout or {}
        p_feat, p_seq, p_enc, seed = dr.get("p_feat", 0.0), dr.get("p_seq", 0.0), dr.get("p_enc", 0.0), dr.get("seed", 0)
        feat_mask = {}
        if p_feat > 0.0:                                  # feat_drop: flat index over the [B, F, dim] input
            if vis is not None:
                feat_mask["visual"] = drop_mask(np.arange(np.size(vis)).reshape(np.shape(vis)), p_feat, seed, DROP_FEAT_VIS).astype(dt)
                vis = np.asarray(vis, dt) * feat_mask["visual"]
            if aud is not None:
                feat_mask["audio"] = drop_mask(np.arange(np.size(aud)).reshape(np.shape(aud)), p_feat, seed, DROP_FEAT_AUD).astype(dt)
                aud = np.asarray(aud, dt) * feat_mask["audio"]

        def acc(key, val):
            g[key] = g[key] + val if key in g else val

        # ------------------------------------------------------------------ forward, keeping what the backward needs
        t = np.asarray(times, dt)
        t_pre, t_act = [], [t]
        for i in (0, 2, 4):
            t_pre.append(t_act[-1] @ s[f"time_mlp.{i}.weight"].T + s[f"time_mlp.{i}.bias"])
            t_act.append(_relu(t_pre[-1]))
        te, te_ln = _ln_fwd(t_act[-1], s["time_mlp.6.weight"], s["time_mlp.6.bias"])

        emb = {}
        for which, x in (("visual", vis), ("audio", aud)):
            if x is None or not getattr(cfg, f"has_{which}_input"):
                continue
            p = f"feature_encoding.{which}_embedder."
            xin = np.asarray(x, dt)
            pre = xin @ s[p + "1.weight"].T + s[p + "1.bias"]
            out, ln = _ln_fwd(_gelu(pre), s[p + "3.weight"], s[p + "3.bias"])
            emb[which] = (xin, pre, ln, out)

        x = self.assemble(vis, aud, te, Qv, Qa)                       # same code path as the forward oracle
        B, S, E = x.shape
        H, hd, d = cfg.nhead, cfg.head_dim, cfg.d_model
        mask = self.mask(S)
        Ft = cfg.F_tot

        def row_mask(width, p, site, layer=0):
            return row_drop_mask(B, S, Ft, width, p, seed, site, layer).astype(dt)

        def attn_mask(layer):
            return attn_drop_mask(B, H, S, Ft, p_enc, seed, layer).astype(dt)

        seq_m = None
        if p_seq > 0.0:
            seq_m = row_mask(E, p_seq, DROP_SEQ)
            x = x * seq_m
        layers = []
        for l in range(cfg.num_layers):
            p = f"{cfg.encoder_prefix}.layers.{l}."
            qkv = x @ s[p + "self_attn.in_proj_weight"].T + s[p + "self_attn.in_proj_bias"]
            q = (qkv[..., :E] * dt.type(hd ** -0.5)).reshape(B, S, H, hd).transpose(0, 2, 1, 3)
            k = qkv[..., E:2 * E].reshape(B, S, H, hd).transpose(0, 2, 1, 3)
            v = qkv[..., 2 * E:].reshape(B, S, H, hd).transpose(0, 2, 1, 3)
            sc = np.where(mask[None, None], dt.type(-np.inf), q @ k.transpose(0, 1, 3, 2))
            pr = np.exp(sc - sc.max(axis=-1, keepdims=True))
            pr = pr / pr.sum(axis=-1, keepdims=True)
            dm = None
            if p_enc > 0.0:
                dm = (attn_mask(l), row_mask(E, p_enc, DROP_SUB1, l), row_mask(cfg.FF, p_enc, DROP_FFN, l), row_mask(E, p_enc, DROP_SUB2, l))
            prd = pr * dm[0] if dm else pr                            # the probabilities that meet V (F.multi_head_attention_forward's dropout)
            ctx = (prd @ v).transpose(0, 2, 1, 3).reshape(B, S, E)
            a = ctx @ s[p + "self_attn.out_proj.weight"].T + s[p + "self_attn.out_proj.bias"]
            if dm:
                a = a * dm[1]
            x1, ln1 = _ln_fwd(x + a, s[p + "norm1.weight"], s[p + "norm1.bias"])
            hpre = x1 @ s[p + "linear1.weight"].T + s[p + "linear1.bias"]
            hact = _gelu(hpre)
            if dm:
                hact = hact * dm[2]
            f = hact @ s[p + "linear2.weight"].T + s[p + "linear2.bias"]
            if dm:
                f = f * dm[3]
            x2, ln2 = _ln_fwd(x1 + f, s[p + "norm2.weight"], s[p + "norm2.bias"])
            layers.append((x, q, k, v, pr, prd, dm, ctx, ln1, x1, hpre, hact, ln2))
            x = x2
        out = self.heads(x, Qv, Qa)
        out["feats"] = x[:, :cfg.F_tot]
        out["time_encodings"] = te

        # ------------------------------------------------------------------ backward
        dx = np.zeros_like(x)
        if cot.get("feats") is not None:
            dx[:, :cfg.F_tot] += np.asarray(cot["feats"], dt)
        hc = cfg.head_classes()
        has_v, has_a = "visual" in cfg.data_modality, "audio" in cfg.data_modality
        aud_start = S - Qa if (has_a and Qa > 0) else S

        def fc_bwd(name, lo, hi, key):
            if cot.get(key) is None or out.get(key) is None:
                return
            if hi == lo:                                  # an empty slice (Qa = 0): autograd still leaves zero gradients
                acc(f"cls_head.{name}.weight", np.zeros_like(s[f"cls_head.{name}.weight"]))
                acc(f"cls_head.{name}.bias", np.zeros_like(s[f"cls_head.{name}.bias"]))
                return
            dy = np.asarray(cot[key], dt).reshape(B, hi - lo, -1)
            ddx, dw, db = _lin_bwd(dy, x[:, lo:hi], s[f"cls_head.{name}.weight"])
            dx[:, lo:hi] += ddx
            acc(f"cls_head.{name}.weight", dw); acc(f"cls_head.{name}.bias", db)

        def reg_bwd(name, lo, hi, key):
            if cot.get(key) is None or out.get(key) is None:
                return
            p = f"reg_head.{name}."
            rows = x[:, lo:hi]
            p0 = rows @ s[p + "0.weight"].T + s[p + "0.bias"]; a0 = _relu(p0)
            p2 = a0 @ s[p + "2.weight"].T + s[p + "2.bias"]; a2 = _relu(p2)
            y = out[key].reshape(B, hi - lo, 2)
            dy = np.asarray(cot[key], dt).reshape(B, hi - lo, 2) * y * (1.0 - y)
            d2, dw, db = _lin_bwd(dy, a2, s[p + "4.weight"]); acc(p + "4.weight", dw); acc(p + "4.bias", db)
            d2 = d2 * (p2 > 0)
            d0, dw, db = _lin_bwd(d2, a0, s[p + "2.weight"]); acc(p + "2.weight", dw); acc(p + "2.bias", db)
            d0 = d0 * (p0 > 0)
            ddx, dw, db = _lin_bwd(d0, rows, s[p + "0.weight"]); acc(p + "0.weight", dw); acc(p + "0.bias", db)
            dx[:, lo:hi] += ddx

        if cfg.variant == RECOGNITION:
            act_start = aud_start - Qv
            if has_v:
                if hc["verb"]:
                    fc_bwd("fc_visual_verb", act_start - 2 * Qv, act_start - Qv, "verb")
                    fc_bwd("fc_visual_noun", act_start - Qv, act_start, "noun")
                fc_bwd("fc_visual_action", act_start, aud_start, "action")
            if has_a:
                fc_bwd("fc_audio_action", aud_start, S, "audio")
        else:
            vis_start = aud_start - Qv
            if has_v:
                if hc["verb"]:
                    fc_bwd("fc_visual_verb", vis_start, aud_start, "verb")
                    fc_bwd("fc_visual_noun", vis_start, aud_start, "noun")
                fc_bwd("fc_visual_action", vis_start, aud_start, "action")
                reg_bwd("fc_visual_action", vis_start, aud_start, "reg_v")
            if has_a:
                fc_bwd("fc_audio_action", aud_start, S, "audio")
                reg_bwd("fc_audio_action", aud_start, S, "reg_a")

        for l in reversed(range(cfg.num_layers)):
            p = f"{cfg.encoder_prefix}.layers.{l}."
            xin, q, k, v, pr, prd, dm, ctx, ln1, x1, hpre, hact, ln2 = layers[l]
            dz2, dw, db = _ln_bwd(dx, ln2, s[p + "norm2.weight"]); acc(p + "norm2.weight", dw); acc(p + "norm2.bias", db)
            df = dz2 * dm[3] if dm else dz2
            dh, dw, db = _lin_bwd(df, hact, s[p + "linear2.weight"]); acc(p + "linear2.weight", dw); acc(p + "linear2.bias", db)
            if dm:
                dh = dh * dm[2]
            dhp = _gelu_bwd(dh, hpre)
            dx1, dw, db = _lin_bwd(dhp, x1, s[p + "linear1.weight"]); acc(p + "linear1.weight", dw); acc(p + "linear1.bias", db)
            dx1 = dx1 + dz2
            dz1, dw, db = _ln_bwd(dx1, ln1, s[p + "norm1.weight"]); acc(p + "norm1.weight", dw); acc(p + "norm1.bias", db)
            da = dz1 * dm[1] if dm else dz1
            dctx, dw, db = _lin_bwd(da, ctx, s[p + "self_attn.out_proj.weight"])
            acc(p + "self_attn.out_proj.weight", dw); acc(p + "self_attn.out_proj.bias", db)
            dctx = dctx.reshape(B, S, H, hd).transpose(0, 2, 1, 3)
            dpr = dctx @ v.transpose(0, 1, 3, 2)
            if dm:
                dpr = dpr * dm[0]
            dv = prd.transpose(0, 1, 3, 2) @ dctx
            dsc = pr * (dpr - (dpr * pr).sum(axis=-1, keepdims=True))          # masked entries have pr = 0
            dq = (dsc @ k) * dt.type(hd ** -0.5)
            dk = dsc.transpose(0, 1, 3, 2) @ q

            def merge(tq):
                return tq.transpose(0, 2, 1, 3).reshape(B, S, E)
            dqkv = np.concatenate([merge(dq), merge(dk), merge(dv)], axis=-1)
            dxa, dw, db = _lin_bwd(dqkv, xin, s[p + "self_attn.in_proj_weight"])
            acc(p + "self_attn.in_proj_weight", dw); acc(p + "self_attn.in_proj_bias", db)
            dx = dz1 + dxa

        # ------------------------------------------------------------------ token assembly (encodings.py) backward
        if seq_m is not None:
            dx = dx * seq_m
        dte = np.zeros_like(te)
        if cot.get("time_encodings") is not None:
            dte += np.asarray(cot["time_encodings"], dt)
        F = cfg.num_feats
        fe = "feature_encoding."
        demb = {}

        def cls_bwd(param, rows, te_lo, te_hi, mod_key):
            blk = dx[:, rows[0]:rows[1]]
            acc(fe + param, blk[..., :d].sum(axis=(0, 1)).reshape(1, 1, d))
            dte[:, te_lo:te_hi] += blk[..., d:]
            if mod_key:
                acc(fe + mod_key, blk.sum(axis=(0, 1)).reshape(1, 1, E))

        T_ = te.shape[1]
        if cfg.input_modality == "audio_visual":
            for which, lo, mk in (("visual", 0, "visual_modality_encoding"), ("audio", F, "audio_modality_encoding")):
                blk = dx[:, lo:lo + F]
                demb[which] = blk[..., :d]
                dte[:, lo:lo + F] += blk[..., d:]
                acc(fe + mk, blk.sum(axis=(0, 1)).reshape(1, 1, E))
            row = 2 * F
            if has_v and Qv > 0:
                names = (["visual_verb_cls", "visual_noun_cls"] if cfg.verb_noun_tokens else []) + ["visual_action_cls"]
                for nm in names:
                    cls_bwd(nm, (row, row + Qv), 2 * F, 2 * F + Qv, "visual_modality_encoding")
                    row += Qv
            if has_a and Qa > 0:
                cls_bwd("audio_action_cls", (row, row + Qa), T_ - Qa, T_, "audio_modality_encoding")
        elif cfg.input_modality == "visual":
            blk = dx[:, :F]
            demb["visual"] = blk[..., :d]
            dte[:, :F] += blk[..., d:]
            row = F
            if cfg.variant == RECOGNITION:
                names = (["verb_cls", "noun_cls"] if cfg.include_verb_noun else []) + ["action_cls"]
            else:
                names = ["visual_action_cls"]
            for nm in names:
                cls_bwd(nm, (row, row + Qv), F, F + Qv, None)
                row += Qv
        else:
            blk = dx[:, :F]
            demb["audio"] = blk[..., :d]
            dte[:, :F] += blk[..., d:]
            cls_bwd("action_cls" if cfg.variant == RECOGNITION else "audio_action_cls", (F, F + Qa), F, F + Qa, None)

        for which, de in demb.items():
            xin, pre, ln, _ = emb[which]
            p = f"feature_encoding.{which}_embedder."
            dact, dw, db = _ln_bwd(de, ln, s[p + "3.weight"]); acc(p + "3.weight", dw); acc(p + "3.bias", db)
            dpre = _gelu_bwd(dact, pre)
            dxin, dw, db = _lin_bwd(dpre, xin, s[p + "1.weight"]); acc(p + "1.weight", dw); acc(p + "1.bias", db)
            if which in feat_mask:
                dxin = dxin * feat_mask[which]
            g["input.vis" if which == "visual" else "input.aud"] = dxin

        # ------------------------------------------------------------------ time MLP backward (tim.py:66-74)
        dcur, dw, db = _ln_bwd(dte, te_ln, s["time_mlp.6.weight"]); acc("time_mlp.6.weight", dw); acc("time_mlp.6.bias", db)
        for j, i in reversed(list(enumerate((0, 2, 4)))):
            dcur = dcur * (t_pre[j] > 0)
            dcur, dw, db = _lin_bwd(dcur, t_act[j], s[f"time_mlp.{i}.weight"])
            acc(f"time_mlp.{i}.weight", dw); acc(f"time_mlp.{i}.bias", db)
        return out, g
