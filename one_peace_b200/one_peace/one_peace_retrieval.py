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


@register_model("one_peace_retrieval", dataclass=OnePeaceRetrievalConfig)
class OnePeaceRetrievalModel(OnePeaceBaseModel):
    def __init__(self, cfg: OnePeaceRetrievalConfig, src_dict, head_type):
        super().__init__(cfg, src_dict)
        embed_dim = self.cfg.encoder.embed_dim
        self.head_type = head_type
        cfg.encoder.use_text_moe = head_type in ("text", "vl", "al", "val")
        cfg.encoder.use_image_moe = head_type in ("image", "vl", "val")
        cfg.encoder.use_audio_moe = head_type in ("audio", "al", "val")
        self.encoder_wrapper = ModelWrapper(cfg.encoder, src_dict, use_text_norm=cfg.encoder.use_text_moe,
                                            use_image_norm=cfg.encoder.use_image_moe,
                                            use_audio_norm=cfg.encoder.use_audio_moe,
                                            num_layers=cfg.encoder.layers if cfg.copy_rel_pos_table else None)
        if cfg.encoder.use_text_moe:
            self.text_proj = Linear(embed_dim, embed_dim)
        if cfg.encoder.use_image_moe:
            self.image_proj = Linear(embed_dim, embed_dim)
        if cfg.encoder.use_audio_moe:
            self.audio_proj = Linear(embed_dim, embed_dim)
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
        super().upgrade_state_dict_named(state_dict, name)
        self.remove_pretraining_modules(state_dict)
        prefix = name + "." if name != "" else ""
        for param_name, _ in self.state_dict().items():
            if (prefix + param_name) not in state_dict:
                state_dict[prefix + param_name] = self.state_dict()[param_name]

    def remove_pretraining_modules(self, state_dict):
        for param_name in list(state_dict.keys()):
            if self.head_type not in ("text", "vl", "al", "val") and "text_" in param_name:
                del state_dict[param_name]
            elif self.head_type not in ("image", "vl", "val") and "image_" in param_name:
                del state_dict[param_name]
            elif self.head_type not in ("audio", "al", "val") and "audio_" in param_name:
                del state_dict[param_name]
