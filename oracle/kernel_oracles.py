"""Kernel-boundary oracles in the library's own data layout - TEST INFRASTRUCTURE ONLY (groundwork for the training leg).

The end-to-end oracles (tim_oracle.py, tim_oracle_bwd.py) restate the reference's dense [B, S, E] formulation. The CUDA kernels work on
the two-stream token buffer (rows [0, B*Ft) feature tokens, rows [B*Ft, B*Ft + B*Qt) query tokens) and exploit the structure of the
reference's mask (recognition/.../models/tim.py:161-166: every token attends to the Ft feature keys of its clip, a query token
additionally to its own key). These functions state forward and backward of the attention core AT THAT BOUNDARY, mask-aware, so
that a backward kernel can be checked in isolation the way tim_test_attention checks the forward one. tests/test_oracle_golden.py
checks them against the dense masked formulation (the reference's), forward and gradient.

Conventions of the kernels: qkv is [M, 3E] = [q | k | v] per row, head h owns columns h*hd .. (h+1)*hd of each third; the softmax
is exp2 of the scores (head_dim**-0.5 * log2(e) is folded into the q columns by the in_proj weights); out is [M, E].
"""
from __future__ import annotations

import numpy as np

LN2 = float(np.log(2.0))


def _split(qkv, B, Ft, Qt, H, hd):
    E = H * hd
    f = qkv[:B * Ft].reshape(B, Ft, 3, H, hd)
    q = qkv[B * Ft:].reshape(B, Qt, 3, H, hd)
    return f, q, E


def attention_fwd_two_stream(qkv, B, Ft, Qt, H, hd):
    """-> (out [M, E], cache). Feature rows: softmax over the Ft feature keys. Query rows: softmax over Ft feature keys + own key."""
    qkv = np.asarray(qkv)
    f, q, E = _split(qkv, B, Ft, Qt, H, hd)
    kf, vf = f[:, :, 1], f[:, :, 2]                                   # [B, Ft, H, hd]
    out = np.empty((qkv.shape[0], E), qkv.dtype)
    # feature rows
    sf = np.einsum("bihd,bjhd->bhij", f[:, :, 0], kf) * LN2             # natural-log scores
    pf = np.exp(sf - sf.max(-1, keepdims=True))
    pf /= pf.sum(-1, keepdims=True)
    out[:B * Ft] = np.einsum("bhij,bjhd->bihd", pf, vf).reshape(B * Ft, E)
    # query rows: Ft feature keys + the row's own key / value
    if Qt:
        sq = np.einsum("bihd,bjhd->bhij", q[:, :, 0], kf) * LN2         # [B, H, Qt, Ft]
        ss = np.einsum("bihd,bihd->bhi", q[:, :, 0], q[:, :, 1]) * LN2   # [B, H, Qt] own-key score
        m = np.maximum(sq.max(-1), ss)
        ef, es = np.exp(sq - m[..., None]), np.exp(ss - m)
        l = ef.sum(-1) + es
        pq, ps = ef / l[..., None], es / l
        oq = np.einsum("bhij,bjhd->bihd", pq, vf) + ps.transpose(0, 2, 1)[..., None] * q[:, :, 2]
        out[B * Ft:] = oq.reshape(B * Qt, E)
    else:
        pq = ps = None
    return out, (pf, pq, ps)


def attention_bwd_two_stream(qkv, dout, B, Ft, Qt, H, hd):
    """d L / d qkv [M, 3E] for L = <out, dout>, by the same structure: K_f / V_f gradients accumulate over the feature rows AND every
    query row of the clip; the own-key / own-value terms touch only the query row itself."""
    qkv, dout = np.asarray(qkv), np.asarray(dout)
    f, q, E = _split(qkv, B, Ft, Qt, H, hd)
    _, (pf, pq, ps) = attention_fwd_two_stream(qkv, B, Ft, Qt, H, hd)
    kf, vf = f[:, :, 1], f[:, :, 2]
    dqkv = np.zeros_like(qkv)
    df = dqkv[:B * Ft].reshape(B, Ft, 3, H, hd)
    dq_ = dqkv[B * Ft:].reshape(B, Qt, 3, H, hd)
    dof = dout[:B * Ft].reshape(B, Ft, H, hd)
    # feature rows
    dpf = np.einsum("bihd,bjhd->bhij", dof, vf)
    dsf = pf * (dpf - (dpf * pf).sum(-1, keepdims=True)) * LN2
    df[:, :, 0] += np.einsum("bhij,bjhd->bihd", dsf, kf)
    df[:, :, 1] += np.einsum("bhij,bihd->bjhd", dsf, f[:, :, 0])
    df[:, :, 2] += np.einsum("bhij,bihd->bjhd", pf, dof)
    if Qt:
        doq = dout[B * Ft:].reshape(B, Qt, H, hd)
        dpq = np.einsum("bihd,bjhd->bhij", doq, vf)                                   # [B, H, Qt, Ft]
        dps = np.einsum("bihd,bihd->bhi", doq, q[:, :, 2])                            # [B, H, Qt]
        dot = (dpq * pq).sum(-1) + dps * ps                                           # sum_j p_j dp_j over all Ft + 1 keys
        dsq = pq * (dpq - dot[..., None]) * LN2
        dss = ps * (dps - dot) * LN2
        dq_[:, :, 0] += np.einsum("bhij,bjhd->bihd", dsq, kf) + dss.transpose(0, 2, 1)[..., None] * q[:, :, 1]
        dq_[:, :, 1] += dss.transpose(0, 2, 1)[..., None] * q[:, :, 0]               # own key
        dq_[:, :, 2] += ps.transpose(0, 2, 1)[..., None] * doq                        # own value
        df[:, :, 1] += np.einsum("bhij,bihd->bjhd", dsq, q[:, :, 0])                   # K_f from the query rows
        df[:, :, 2] += np.einsum("bhij,bihd->bjhd", pq, doq)                          # V_f from the query rows
    return dqkv


def layernorm_bwd_rows(dy, x, gamma, eps=1e-5):
    """Row kernel boundary: y = LayerNorm(x) * gamma + beta over the last axis -> (dx, dgamma, dbeta)."""
    mu = x.mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(((x - mu) ** 2).mean(-1, keepdims=True) + eps)
    xhat = (x - mu) * rstd
    dxhat = dy * gamma
    dx = rstd * (dxhat - dxhat.mean(-1, keepdims=True) - xhat * (dxhat * xhat).mean(-1, keepdims=True))
    return dx, (dy * xhat).reshape(-1, x.shape[-1]).sum(0), dy.reshape(-1, x.shape[-1]).sum(0)
