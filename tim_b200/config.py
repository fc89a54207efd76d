"""Static description of a TIM model: constructor arguments, derived sizes, token layout and
the checkpoint (state_dict) key/shape contract.

Everything here mirrors what the reference builds in
  recognition/time_interval_machine/models/tim.py:17-145   (recognition TIM.__init__/_create_model)
  detection/time_interval_machine/models/tim.py:17-142     (detection TIM)
  */models/helpers/encodings.py, */models/helpers/head.py  (parameter names)
and is pure Python (no torch, no CUDA) so both the host plugin and the tests can use it.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

RECOGNITION = "recognition"
DETECTION = "detection"
_MODALITIES = ("audio_visual", "visual", "audio")

# compute_dtype codes shared with include/tim_b200.h
DTYPE_FP32 = 0
DTYPE_BF16 = 1
DTYPE_FP16 = 2
DTYPE_NAMES = {DTYPE_FP32: "fp32", DTYPE_BF16: "bf16", DTYPE_FP16: "fp16"}
DTYPE_CODES = {v: k for k, v in DTYPE_NAMES.items()}


def _is_listlike(x) -> bool:
    return isinstance(x, (list, tuple))


@dataclass
class TIMConfig:
    """Constructor arguments of the reference TIM (same names, same defaults)."""
    num_class: Any = field(default_factory=lambda: [[97, 300, 3806], 44])
    visual_input_dim: int = 1024
    audio_input_dim: int = 2304
    d_model: int = 512
    feedforward_scale: int = 4
    nhead: int = 8
    num_layers: int = 6
    input_modality: str = "audio_visual"
    data_modality: str = "audio_visual"
    num_feats: int = 50              # per modality, as passed to the constructor
    include_verb_noun: bool = True
    variant: str = RECOGNITION

    def __post_init__(self):
        if self.variant not in (RECOGNITION, DETECTION):
            raise ValueError(f"variant must be recognition|detection, got {self.variant!r}")
        if self.input_modality not in _MODALITIES or self.data_modality not in _MODALITIES:
            raise ValueError("modalities must be one of %s" % (_MODALITIES,))
        if self.input_modality != "audio_visual" and self.data_modality != self.input_modality:
            # the uni-modal encodings only build their own CLS tokens (encodings.py:7-121)
            raise ValueError("uni-modal input requires data_modality == input_modality")
        if (2 * self.d_model) % self.nhead:
            raise ValueError("2*d_model must be divisible by nhead")

    # ---- derived sizes (tim.py:46, 115-121: width is 2*d_model, FF = d_model*scale) ----
    @property
    def E(self) -> int:
        return 2 * self.d_model

    @property
    def FF(self) -> int:
        return self.d_model * self.feedforward_scale

    @property
    def head_dim(self) -> int:
        return self.E // self.nhead

    @property
    def has_visual_input(self) -> bool:
        return self.input_modality in ("audio_visual", "visual")

    @property
    def has_audio_input(self) -> bool:
        return self.input_modality in ("audio_visual", "audio")

    @property
    def F_tot(self) -> int:
        """Feature tokens per clip (tim.py:87 doubles num_feats for audio_visual)."""
        return self.num_feats * (2 if self.input_modality == "audio_visual" else 1)

    @property
    def has_reg_head(self) -> bool:
        return self.variant == DETECTION

    @property
    def encoder_prefix(self) -> str:
        return "transformer_encoder" if self.variant == RECOGNITION else "backbone"

    # ---- classification head widths, following head.py's own isinstance() logic ----
    def head_classes(self) -> Dict[str, int]:
        """{'verb': n, 'noun': n, 'action': n, 'audio': n}; 0 = head absent.

        recognition head.py:4-81: AudioVisualCLSHead tests isinstance(num_class[0], list);
        detection head.py:7-93: it tests isinstance(num_class, list) on the whole argument.
        """
        out = {"verb": 0, "noun": 0, "action": 0, "audio": 0}
        nc = self.num_class
        if self.data_modality == "audio_visual":
            if self.variant == RECOGNITION:
                vn = isinstance(nc[0], list)
            else:
                vn = isinstance(nc, list) and _is_listlike(nc[0])
            if vn:
                out["verb"], out["noun"], out["action"] = (int(v) for v in nc[0])
            else:
                out["action"] = int(nc[0])
            out["audio"] = int(nc[1])
        elif self.data_modality == "visual":
            v = nc[0]
            if isinstance(v, list):
                out["verb"], out["noun"], out["action"] = (int(x) for x in v)
            else:
                out["action"] = int(v)
        else:
            out["audio"] = int(nc[1])
        return out

    @property
    def verb_noun_tokens(self) -> bool:
        """True when verb and noun CLS tokens exist as separate token groups
        (recognition encodings.py:30-35,166-171). Detection never builds them."""
        return (self.variant == RECOGNITION and self.include_verb_noun
                and "visual" in self.data_modality)

    # ---- token layout ----
    def query_tokens(self, Qv: int, Qa: int) -> int:
        n = 0
        if "visual" in self.data_modality and Qv > 0:
            n += Qv * (3 if self.verb_noun_tokens else 1)
        if "audio" in self.data_modality and Qa > 0:
            n += Qa
        return n

    def seq_len(self, Qv: int, Qa: int) -> int:
        return self.F_tot + self.query_tokens(Qv, Qa)

    def flops_fwd_per_clip(self, Qv: int, Qa: int) -> float:
        """Algorithmic (mask-aware) forward FLOPs per clip, SURVEY.md §8(d) formula."""
        d, E, FF, L = self.d_model, self.E, self.FF, self.num_layers
        F_tot = self.F_tot
        Qt = self.query_tokens(Qv, Qa)
        S = F_tot + Qt
        T = F_tot + (Qv if "visual" in self.data_modality else 0) + (Qa if "audio" in self.data_modality else 0)
        fl = T * (4 * d + 4 * d * d)
        if self.has_visual_input:
            fl += self.num_feats * 2 * self.visual_input_dim * d
        if self.has_audio_input:
            fl += self.num_feats * 2 * self.audio_input_dim * d
        fl += L * S * (8 * E * E + 4 * E * FF)
        fl += L * 4 * E * (F_tot * F_tot + Qt * (F_tot + 1))
        hc = self.head_classes()
        if "visual" in self.data_modality:
            fl += Qv * 2 * E * (hc["verb"] + hc["noun"] + hc["action"])
            if self.has_reg_head:
                fl += Qv * (E * E + E * E // 2 + 2 * E)
        if "audio" in self.data_modality:
            fl += Qa * 2 * E * hc["audio"]
            if self.has_reg_head:
                fl += Qa * (E * E + E * E // 2 + 2 * E)
        return float(fl)


def state_dict_spec(cfg: TIMConfig, include_drloc: bool = True) -> "OrderedDict[str, Tuple[int, ...]]":
    """Checkpoint contract: every parameter name and shape the reference model registers
    (verified against the imported reference in tools/make_golden.py)."""
    d, E, FF = cfg.d_model, cfg.E, cfg.FF
    sd: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for i, (o, k) in zip((0, 2, 4), ((d, 2), (d, d), (d, d))):
        sd[f"time_mlp.{i}.weight"] = (o, k)
        sd[f"time_mlp.{i}.bias"] = (o,)
    sd["time_mlp.6.weight"] = (d,)
    sd["time_mlp.6.bias"] = (d,)

    fe = "feature_encoding."
    if cfg.input_modality == "audio_visual":
        sd[fe + "visual_modality_encoding"] = (1, 1, E)
        sd[fe + "audio_modality_encoding"] = (1, 1, E)
        if "visual" in cfg.data_modality:
            sd[fe + "visual_action_cls"] = (1, 1, d)
            if cfg.verb_noun_tokens:
                sd[fe + "visual_verb_cls"] = (1, 1, d)
                sd[fe + "visual_noun_cls"] = (1, 1, d)
        if "audio" in cfg.data_modality:
            sd[fe + "audio_action_cls"] = (1, 1, d)
    elif cfg.input_modality == "visual":
        if cfg.variant == RECOGNITION:
            sd[fe + "action_cls"] = (1, 1, d)
            if cfg.include_verb_noun:
                sd[fe + "verb_cls"] = (1, 1, d)
                sd[fe + "noun_cls"] = (1, 1, d)
        else:
            sd[fe + "visual_action_cls"] = (1, 1, d)
    else:
        sd[fe + ("action_cls" if cfg.variant == RECOGNITION else "audio_action_cls")] = (1, 1, d)
    if cfg.has_visual_input:
        sd[fe + "visual_embedder.1.weight"] = (d, cfg.visual_input_dim)
        sd[fe + "visual_embedder.1.bias"] = (d,)
        sd[fe + "visual_embedder.3.weight"] = (d,)
        sd[fe + "visual_embedder.3.bias"] = (d,)
    if cfg.has_audio_input:
        sd[fe + "audio_embedder.1.weight"] = (d, cfg.audio_input_dim)
        sd[fe + "audio_embedder.1.bias"] = (d,)
        sd[fe + "audio_embedder.3.weight"] = (d,)
        sd[fe + "audio_embedder.3.bias"] = (d,)

    hc = cfg.head_classes()
    for name, key in (("verb", "fc_visual_verb"), ("noun", "fc_visual_noun"),
                      ("action", "fc_visual_action"), ("audio", "fc_audio_action")):
        if hc[name]:
            sd[f"cls_head.{key}.weight"] = (hc[name], E)
            sd[f"cls_head.{key}.bias"] = (hc[name],)
    if cfg.has_reg_head:
        for mod, key in (("visual", "fc_visual_action"), ("audio", "fc_audio_action")):
            if mod in cfg.data_modality:
                sd[f"reg_head.{key}.0.weight"] = (E // 2, E)
                sd[f"reg_head.{key}.0.bias"] = (E // 2,)
                sd[f"reg_head.{key}.2.weight"] = (E // 2, E // 2)
                sd[f"reg_head.{key}.2.bias"] = (E // 2,)
                sd[f"reg_head.{key}.4.weight"] = (2, E // 2)
                sd[f"reg_head.{key}.4.bias"] = (2,)

    p = cfg.encoder_prefix
    for l in range(cfg.num_layers):
        b = f"{p}.layers.{l}."
        sd[b + "self_attn.in_proj_weight"] = (3 * E, E)
        sd[b + "self_attn.in_proj_bias"] = (3 * E,)
        sd[b + "self_attn.out_proj.weight"] = (E, E)
        sd[b + "self_attn.out_proj.bias"] = (E,)
        sd[b + "norm1.weight"] = (E,)
        sd[b + "norm1.bias"] = (E,)
        sd[b + "linear1.weight"] = (FF, E)
        sd[b + "linear1.bias"] = (FF,)
        sd[b + "linear2.weight"] = (E, FF)
        sd[b + "linear2.bias"] = (E,)
        sd[b + "norm2.weight"] = (E,)
        sd[b + "norm2.bias"] = (E,)
    if include_drloc:
        sd["drloc_mlp.0.weight"] = (d, 4 * d)
        sd["drloc_mlp.0.bias"] = (d,)
        sd["drloc_mlp.2.weight"] = (d, d)
        sd["drloc_mlp.2.bias"] = (d,)
        sd["drloc_mlp.4.weight"] = (1, d)
        sd["drloc_mlp.4.bias"] = (1,)
    return sd


def hot_path_keys(cfg: TIMConfig) -> List[str]:
    """state_dict keys the B200 path consumes (everything except drloc_mlp / pool)."""
    return [k for k in state_dict_spec(cfg, include_drloc=False)]


# ---------------------------------------------------------------------------------------
# The named workloads of BASELINE.json / SURVEY.md §8(d)
# ---------------------------------------------------------------------------------------
def named_config(name: str) -> Tuple[TIMConfig, int, int]:
    """Returns (cfg, Qv, Qa) for 'cfg1' .. 'cfg4' (SURVEY.md §8d table)."""
    if name == "cfg1":   # plumbing: recog L=1 d=512 F=25+25 Qv=Qa=5
        return TIMConfig(num_class=[[97, 300, 3806], 44], num_layers=1, num_feats=25), 5, 5
    if name == "cfg2":   # EPIC-100 recognition, 6L d=512, 50+50 feats, 25+25 queries (100 query tokens)
        return TIMConfig(num_class=[[97, 300, 3806], 44], num_layers=6, num_feats=50), 25, 25
    if name == "cfg3":   # Perception-Test: d=768, 64+64 feats, 200+200 queries, action-only heads (63, 17)
        return TIMConfig(num_class=[63, 17], d_model=768, num_layers=6, num_feats=64,
                         include_verb_noun=False), 200, 200
    if name == "cfg4":   # detection dense queries: visual data modality, Dv=2048, 2048 interval queries
        return TIMConfig(num_class=[97, 44], visual_input_dim=2048, num_layers=6, num_feats=50,
                         data_modality="visual", include_verb_noun=False, variant=DETECTION), 2048, 0
    raise KeyError(name)
