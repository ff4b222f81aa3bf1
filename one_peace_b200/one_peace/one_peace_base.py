"""Drop-in for models/one_peace/one_peace_base.py: ``ModelWrapper`` (adapters + shared encoder + feature
slicing, :39-129), ``OnePeaceBaseModel`` (:238-259) and ``init_one_peace_params`` (:262-274)."""
from typing import Optional

import torch
import torch.nn as nn

from .. import kernels as K
from ..adapter.audio import AudioAdapter
from ..adapter.image import ImageAdapter
from ..adapter.text import TextAdapter
from ..components import trunc_normal_
from ..fairseq_compat import BaseFairseqModel, register_model
from ..transformer.transformer_encoder import TransformerEncoder
from ..unify_model_config import UnifyModelConfig


class ModelWrapper(nn.Module):
    def __init__(self, cfg, src_dict=None, use_text_norm=True, use_image_norm=True, use_audio_norm=True,
                 num_layers=None):
        super().__init__()
        embed_dim, heads = cfg.embed_dim, cfg.attention_heads
        if cfg.use_text_moe:
            self.text_adapter = TextAdapter(cfg.text_adapter, embed_dim, heads, src_dict, num_layers)
        if cfg.use_image_moe:
            self.image_adapter = ImageAdapter(cfg.image_adapter, embed_dim, heads, num_layers)
        if cfg.use_audio_moe:
            self.audio_adapter = AudioAdapter(cfg.audio_adapter, embed_dim, heads, num_layers)
        self.fusion_model = TransformerEncoder(cfg, src_dict, use_text_norm=use_text_norm,
                                               use_image_norm=use_image_norm, use_audio_norm=use_audio_norm)

    def adapt(self, encoder_type, src_tokens=None, src_images=None, src_audios=None, audio_padding_masks=None):
        if encoder_type == "text":
            return self.text_adapter(src_tokens)
        if encoder_type == "image":
            return self.image_adapter(src_images)
        if encoder_type == "audio":
            return self.audio_adapter(src_audios, audio_padding_masks)
        raise NotImplementedError(f"encoder_type={encoder_type!r}")

    def encode_cls(self, encoder_type, **inputs):
        """Fast path of the embedding API: adapter -> 40 layers -> final LayerNorm of the CLS rows only
        (the retrieval heads read x[:, 0, :], one_peace_retrieval.py:107-119).  Returns bf16 [B, d]."""
        info = self.adapt(encoder_type, **inputs)
        x, _ = self.fusion_model.run_layers(info, encoder_type)
        B, S, d = x.shape
        cls = torch.empty(B, d, dtype=torch.bfloat16, device=x.device)
        pk = self.fusion_model.final_norm_pack(encoder_type)
        if pk is None:
            raise NotImplementedError("retrieval heads always build the modality layer norm")
        K.layernorm(x, pk[0], pk[1], cls, rows=B, dim=d, ld_in=S * d, ld_out=d, eps=pk[2])
        return cls

    def forward(self, src_tokens: Optional[torch.Tensor] = None, text_preserve_ids=None, text_preserve_embed=None,
                text_mask_token=None, src_images: Optional[torch.Tensor] = None, image_preserve_ids=None,
                image_preserve_embed=None, image_mask_token=None, is_second_image: bool = False,
                src_audios: Optional[torch.Tensor] = None, audio_padding_masks: Optional[torch.Tensor] = None,
                audio_preserve_ids=None, audio_preserve_embed=None, audio_mask_token=None,
                encoder_type: Optional[str] = None, return_padding_mask: bool = False):
        """Reference signature (one_peace_base.py:68-129); returns per-token features (B,S,d) fp32."""
        general = any(v is not None for v in (text_preserve_ids, text_preserve_embed, image_preserve_ids, image_preserve_embed,
                                              audio_preserve_ids, audio_preserve_embed)) or encoder_type in ("vl", "al") or \
            (torch.is_grad_enabled() and any(q.requires_grad for q in self.parameters()))
        if general:
            return self.forward_general(src_tokens, text_preserve_ids, text_preserve_embed, text_mask_token, src_images,
                                        image_preserve_ids, image_preserve_embed, image_mask_token, src_audios,
                                        audio_padding_masks, audio_preserve_ids, audio_preserve_embed, audio_mask_token,
                                        encoder_type, return_padding_mask)
        info = self.adapt(encoder_type, src_tokens=src_tokens, src_images=src_images, src_audios=src_audios,
                          audio_padding_masks=audio_padding_masks)
        infos = {"text": None, "image": None, "audio": None}
        infos[encoder_type] = info
        out = self.fusion_model(infos["text"], infos["image"], infos["audio"], encoder_type=encoder_type)
        feats = out["encoder_out"][0].transpose(0, 1)
        pad = out["encoder_padding_mask"]
        res = [None, None, None]
        pads = [None, None, None]
        i = ("text", "image", "audio").index(encoder_type)
        res[i], pads[i] = feats, pad
        return (*res, *pads) if return_padding_mask else tuple(res)


    def forward_general(self, src_tokens, text_preserve_ids, text_preserve_embed, text_mask_token, src_images,
                        image_preserve_ids, image_preserve_embed, image_mask_token, src_audios, audio_padding_masks,
                        audio_preserve_ids, audio_preserve_embed, audio_mask_token, encoder_type, return_padding_mask):
        """one_peace_base.py:68-129 for every case the single-modality inference path does not cover: concatenated 'vl' / 'al'
        sequences, preserve_ids student passes, the decoder's mask-token canvas, and per-token features with gradients."""
        if encoder_type not in ("text", "image", "audio", "vl", "al"):
            raise NotImplementedError(f"encoder_type={encoder_type!r}")        # 'val' raises in the reference too (:136-137)
        parts = []
        if encoder_type in ("text", "vl", "al"):
            parts.append(("text",) + tuple(self.text_adapter.embed_general(src_tokens, text_preserve_ids, text_preserve_embed,
                                                                           text_mask_token)))
        if encoder_type in ("image", "vl"):
            parts.append(("image",) + tuple(self.image_adapter.embed_general(src_images, image_preserve_ids,
                                                                             image_preserve_embed, image_mask_token)))
        if encoder_type in ("audio", "al"):
            parts.append(("audio",) + tuple(self.audio_adapter.embed_general(src_audios, audio_padding_masks, audio_preserve_ids,
                                                                             audio_preserve_embed, audio_mask_token)))
        feats, pads = self.fusion_model.forward_general(parts)
        res, pmask = [None, None, None], [None, None, None]
        for (m, x, _, _), f, p in zip(parts, feats, pads):
            i = ("text", "image", "audio").index(m)
            res[i] = f
            pmask[i] = p.bool() if p is not None else torch.zeros(x.shape[:2], dtype=torch.bool, device=x.device)
        return (*res, *pmask) if return_padding_mask else tuple(res)


@register_model("one_peace_base_b200", dataclass=UnifyModelConfig)
class OnePeaceBaseModel(BaseFairseqModel):
    def __init__(self, cfg: UnifyModelConfig, src_dict):
        super().__init__()
        self.cfg = cfg
        self.src_dict = src_dict

    @classmethod
    def build_model(cls, cfg, task):
        return cls(cfg, task.source_dictionary)

    def no_weight_decay(self):
        # one_peace_base.py:251-259 (names only; the reference's missing comma is reproduced on purpose:
        # the adjacent string literals concatenate, which is what layer_decay.get_parameter_groups sees)
        return {
            'encoder_wrapper.text_adapter.embed_positions.weight', 'encoder_wrapper.text_adapter.cls_embedding',
            'encoder_wrapper.image_adapter.pos_embed', 'encoder_wrapper.image_adapter.cls_embedding',
            'encoder_wrapper.audio_adapter.cls_embedding',
            'decoder_wrapper.text_adapter.embed_positions.weight', 'decoder_wrapper.text_adapter.cls_embedding'
            'decoder_wrapper.image_adapter.pos_embed', 'decoder_wrapper.image_adapter.cls_embedding',
            'decoder_wrapper.audio_adapter.embed_positions.weight', 'decoder_wrapper.audio_adapter.cls_embedding'
        }


def init_one_peace_params(module):
    """one_peace_base.py:262-274."""
    if isinstance(module, nn.Linear):
        trunc_normal_(module.weight)
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)
    elif isinstance(module, nn.LayerNorm):
        if module.elementwise_affine:
            nn.init.constant_(module.bias, 0)
            nn.init.constant_(module.weight, 1.0)
    elif isinstance(module, nn.Conv2d):
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)
