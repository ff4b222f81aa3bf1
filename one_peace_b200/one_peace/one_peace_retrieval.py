"""Drop-in for ``OnePeaceRetrievalModel`` (models/one_peace/one_peace_retrieval.py:34-150) — the model behind
``extract_{text,image,audio}_features`` and the retrieval fine-tune criterion.  Registered under the
reference's name ``one_peace_retrieval`` when fairseq is present (swap via ``--user-dir``)."""
import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from .. import kernels as K
from ..components import Linear, PackCache, bf16, f32
from ..fairseq_compat import register_model
from ..unify_model_config import UnifyModelConfig
from .one_peace_base import ModelWrapper, OnePeaceBaseModel, init_one_peace_params


@dataclass
class OnePeaceRetrievalConfig(UnifyModelConfig):
    copy_rel_pos_table: bool = False


# modalities behind every head type (one_peace_retrieval.py:40-46)
_HEAD_MODALITIES = {"text": ("text",), "image": ("image",), "audio": ("audio",), "vl": ("text", "image"),
                    "al": ("text", "audio"), "val": ("text", "image", "audio")}
_ALL_MODALITIES = ("text", "image", "audio")


@register_model("one_peace_retrieval", dataclass=OnePeaceRetrievalConfig)
class OnePeaceRetrievalModel(OnePeaceBaseModel):
    def __init__(self, cfg: OnePeaceRetrievalConfig, src_dict, head_type):
        super().__init__(cfg, src_dict)
        if head_type not in _HEAD_MODALITIES:
            raise ValueError(f"head_type must be one of {sorted(_HEAD_MODALITIES)}, got {head_type!r}")
        self.head_type = head_type
        self.modalities = _HEAD_MODALITIES[head_type]
        enc = cfg.encoder
        for m in _ALL_MODALITIES:
            setattr(enc, f"use_{m}_moe", m in self.modalities)
        self.encoder_wrapper = ModelWrapper(enc, src_dict, num_layers=enc.layers if cfg.copy_rel_pos_table else None,
                                            **{f"use_{m}_norm": m in self.modalities for m in _ALL_MODALITIES})
        for m in _ALL_MODALITIES:                              # registration order = the reference's parameter order
            if m in self.modalities:
                setattr(self, f"{m}_proj", Linear(enc.embed_dim, enc.embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self.apply(init_one_peace_params)
        self._proj_cache = {}

    def set_num_updates(self, num_updates):
        super().set_num_updates(num_updates)
        self.num_updates = num_updates

    def _proj_pack(self, modality):
        proj = getattr(self, f"{modality}_proj")
        cache = self._proj_cache.setdefault(modality, PackCache())
        return cache.get([proj.weight, proj.bias], lambda: (bf16(proj.weight), f32(proj.bias)))

    def forward(self, src_tokens: Optional[torch.Tensor] = None, src_images: Optional[torch.Tensor] = None,
                src_audios: Optional[torch.Tensor] = None, audio_padding_masks: Optional[torch.Tensor] = None,
                return_logit_scale: bool = False, encoder_type: Optional[str] = None):
        if return_logit_scale:
            with torch.no_grad():
                self.logit_scale.clamp_(0, math.log(100))
            return self.logit_scale.exp()
        if encoder_type not in ("text", "image", "audio"):
            raise NotImplementedError
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self.forward_train(encoder_type, src_tokens=src_tokens, src_images=src_images, src_audios=src_audios,
                                      audio_padding_masks=audio_padding_masks)
        cls = self.encoder_wrapper.encode_cls(encoder_type, src_tokens=src_tokens, src_images=src_images,
                                              src_audios=src_audios, audio_padding_masks=audio_padding_masks)
        w, b = self._proj_pack(encoder_type)
        logits = torch.empty(cls.shape[0], w.shape[0], dtype=torch.float32, device=cls.device)
        K.gemm(cls, w, K.EPI_STORE_F32, logits, bias=b)
        out = K.l2_normalize_rows(logits)
        return out.to(getattr(self, f"{encoder_type}_proj").weight.dtype)

    def forward_train(self, encoder_type, **inputs):
        """Training forward: the same kernels recorded as autograd nodes (one_peace_b200/autograd.py)."""
        from ..autograd import HeadFn
        ew = self.encoder_wrapper
        info = ew.adapt(encoder_type, **inputs)
        x, _ = ew.fusion_model.run_layers(info, encoder_type)
        ln = getattr(ew.fusion_model, f"{encoder_type}_layer_norm")
        proj = getattr(self, f"{encoder_type}_proj")
        out = HeadFn.apply(x, ln.weight, ln.bias, proj.weight, proj.bias, ln.eps)
        return out.to(proj.weight.dtype)

    @classmethod
    def build_model(cls, cfg, task):
        cfg.encoder.image_adapter.rel_bucket_size = task.cfg.patch_image_size // 16
        return cls(cfg, task.source_dictionary, task.cfg.head_type)

    def upgrade_state_dict_named(self, state_dict, name):
        """one_peace_retrieval.py:133-150: prune what this head does not use, then let parameters absent from the
        checkpoint (e.g. a freshly added head) keep their initial values so that the strict load succeeds."""
        super().upgrade_state_dict_named(state_dict, name)
        self.remove_pretraining_modules(state_dict)
        prefix = f"{name}." if name else ""
        for key, value in self.state_dict().items():
            state_dict.setdefault(prefix + key, value)

    def remove_pretraining_modules(self, state_dict):
        unused = [f"{m}_" for m in _ALL_MODALITIES if m not in self.modalities]
        for key in [k for k in state_dict if any(tag in k for tag in unused)]:
            del state_dict[key]
