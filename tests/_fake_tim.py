"""Stand-in for the reference TIM nn.Module on machines where /root/reference does not exist (the GPU box).

It has exactly what tim_b200.plugin.patch_model touches on the real class: the constructor attributes
(recognition/time_interval_machine/models/tim.py:37-53), the parameter tree under the reference's state_dict names
(built from tim_b200.config.state_dict_spec, which tools/make_golden.py checks key-by-key against the real model), the dropout
modules under the reference's names (parameter-free, so the state_dict is unchanged), and for
detection the `backbone` / `inference_queries` / `label_queries` members (detection/.../models/tim.py:140-155,186-270).
It has NO forward of its own: if the patch did not take over, calling it fails.
"""
import torch
from torch import nn

from tim_b200.config import DETECTION, TIMConfig, state_dict_spec


class _Node(nn.Module):
    pass


def _attach(root: nn.Module, dotted: str, value: torch.Tensor):
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            mod.add_module(p, _Node())
        mod = getattr(mod, p)
    mod.register_parameter(parts[-1], nn.Parameter(value))


class FakeTIM(nn.Module):
    def __init__(self, cfg: TIMConfig, sd, feat_drop: float = 0.0, seq_drop: float = 0.0, enc_dropout: float = 0.0):
        super().__init__()
        self.feat_drop, self.seq_drop, self.enc_dropout = feat_drop, seq_drop, enc_dropout
        self.input_modality, self.data_modality = cfg.input_modality, cfg.data_modality
        self.visual_input_dim, self.audio_input_dim = cfg.visual_input_dim, cfg.audio_input_dim
        self.d_model, self.dim_feedforward = cfg.d_model, cfg.d_model * cfg.feedforward_scale
        self.nhead, self.num_layers = cfg.nhead, cfg.num_layers
        self.num_class, self.include_verb_noun = cfg.num_class, cfg.include_verb_noun
        self.num_feats = cfg.F_tot                      # tim.py:87 doubles it for audio_visual input
        self.pool_features = False
        for k, shape in state_dict_spec(cfg).items():
            v = torch.as_tensor(sd[k]).clone().float() if k in sd else torch.zeros(shape)
            assert tuple(v.shape) == tuple(shape), k
            _attach(self, k, v)
        self.feature_encoding.num_feats = cfg.num_feats
        # the reference's dropout modules under their own names (helpers/encodings.py:141,149,177; helpers/transformers.py:73-82)
        fe = self.feature_encoding
        for emb in ("visual_embedder", "audio_embedder"):
            if hasattr(fe, emb):
                getattr(fe, emb).add_module("0", nn.Dropout(feat_drop))
        fe.add_module("dropout", nn.Dropout(seq_drop))
        for layer in getattr(self, cfg.encoder_prefix).layers.children():
            layer.self_attn.dropout = enc_dropout
            for n in ("dropout1", "dropout", "dropout2"):
                layer.add_module(n, nn.Dropout(enc_dropout))
        if cfg.variant == DETECTION:
            assert hasattr(self, "backbone")
            self.iou_threshold = 0.25
            self.label_smoothing = 0.9
            self.inference_queries = torch.zeros(1, 0, 2)

    def label_queries(self, queries, target, modality, thr):           # target prep stays reference code; unused here
        raise AssertionError("label_queries is reference code and is not exercised by these tests")

    def forward(self, *a, **k):
        raise AssertionError("FakeTIM.forward must have been replaced by tim_b200.patch_model")
