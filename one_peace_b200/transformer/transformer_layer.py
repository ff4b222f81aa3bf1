"""Drop-in for ``TransformerEncoderLayer`` / ``GeGLU`` (models/transformer/transformer_layer.py:54-228).

Parameter names match the reference (self_attn.*, self_attn_layer_norm, {text,image,audio}_ffn.{0.wi_0,
0.wi_1,2,3}, final_layer_norm, gamma_1, gamma_2).  One layer forward = 4 tcgen05 GEMMs + 1 attention
kernel + 4 LayerNorm kernels; the residual stream stays fp32 in HBM and is updated in place by the
out_proj / fc2 GEMM epilogues (gamma * (acc + bias) + residual — `fused_dropout_res`, :70-88, eval mode).
"""
import torch
import torch.nn as nn

from .. import kernels as K
from ..components import LayerNorm, Linear, PackCache, bf16, f32
from .multihead_attention import MultiheadAttention


class GeGLU(nn.Module):
    """models/transformer/transformer_layer.py:54-67 — parameter container (wi_0, wi_1: no bias)."""

    def __init__(self, embed_dim, ffn_dim):
        super().__init__()
        self.wi_0 = Linear(embed_dim, ffn_dim, bias=False)
        self.wi_1 = Linear(embed_dim, ffn_dim, bias=False)


def interleave_geglu(w0, w1):
    """[F,d],[F,d] -> bf16 [2F,d] where each 256-row GEMM tile holds 128 rows of wi_0 followed by the matching
    128 rows of wi_1, so the epilogue can form gelu(a) * b inside one accumulator tile."""
    F_, d = w0.shape
    assert F_ % 128 == 0, "ffn_embed_dim must be a multiple of 128"
    return torch.stack([bf16(w0).view(F_ // 128, 128, d), bf16(w1).view(F_ // 128, 128, d)], dim=1).reshape(2 * F_, d).contiguous()


class TransformerEncoderLayer(nn.Module):
    def __init__(self, cfg, drop_path_rate=0.0):
        super().__init__()
        self.cfg = cfg
        self.embed_dim = cfg.embed_dim
        self.ffn_embed_dim = cfg.ffn_embed_dim
        self.self_attn = MultiheadAttention(self.embed_dim, cfg.attention_heads, dropout=cfg.attention_dropout,
                                            scale_heads=cfg.scale_heads, magneto_scale_attn=cfg.magneto_scale_attn)
        self.self_attn_layer_norm = LayerNorm(self.embed_dim)
        self.dropout_prob = cfg.dropout
        self.drop_path_prob = drop_path_rate
        if cfg.use_text_moe:
            self.text_ffn = self.build_geglu_ffn(cfg)
        if cfg.use_image_moe:
            self.image_ffn = self.build_geglu_ffn(cfg)
        if cfg.use_audio_moe:
            self.audio_ffn = self.build_geglu_ffn(cfg)
        self.attn_ln = LayerNorm(self.embed_dim) if cfg.scale_attn else None
        self.final_layer_norm = LayerNorm(self.embed_dim)
        self.gamma_1 = None
        self.gamma_2 = None
        if cfg.use_layer_scale:
            self.gamma_1 = nn.Parameter(cfg.layer_scale_init_value * torch.ones((self.embed_dim)), requires_grad=True)
            self.gamma_2 = nn.Parameter(cfg.layer_scale_init_value * torch.ones((self.embed_dim)), requires_grad=True)
        self._cache = {}

    def build_geglu_ffn(self, cfg):
        # indices 0..3 match the reference Sequential (GeGLU, act-dropout, LayerNorm | Identity, Linear)
        return nn.Sequential(GeGLU(self.embed_dim, self.ffn_embed_dim), nn.Identity(),
                             LayerNorm(self.ffn_embed_dim) if cfg.scale_fc else nn.Identity(),
                             Linear(self.ffn_embed_dim, self.embed_dim))

    def _ffn_pack(self, modality):
        ffn = getattr(self, f"{modality}_ffn")
        cache = self._cache.setdefault(modality, PackCache())
        has_ln = isinstance(ffn[2], nn.LayerNorm)
        ps = [ffn[0].wi_0.weight, ffn[0].wi_1.weight, ffn[3].weight, ffn[3].bias] + ([ffn[2].weight, ffn[2].bias] if has_ln else [])

        def build():
            out = dict(w01=interleave_geglu(ffn[0].wi_0.weight, ffn[0].wi_1.weight), w2=bf16(ffn[3].weight), b2=f32(ffn[3].bias))
            if has_ln:
                out["ln_w"], out["ln_b"] = f32(ffn[2].weight), f32(ffn[2].bias)
            return out
        return cache.get(ps, build)

    def _norm_pack(self):
        cache = self._cache.setdefault("_norm", PackCache())
        ps = [self.self_attn_layer_norm.weight, self.self_attn_layer_norm.bias, self.final_layer_norm.weight,
              self.final_layer_norm.bias] + ([self.gamma_1, self.gamma_2] if self.gamma_1 is not None else [])

        def build():
            out = dict(ln1_w=f32(ps[0]), ln1_b=f32(ps[1]), ln2_w=f32(ps[2]), ln2_b=f32(ps[3]))
            if self.gamma_1 is not None:
                out["g1"], out["g2"] = f32(self.gamma_1), f32(self.gamma_2)
            return out
        return cache.get(ps, build)

    def forward_rows(self, x, bias, key_pad, B, S, modality):
        """x: fp32 [B*S, d] residual stream, updated IN PLACE.  Single-modality sequence
        (encoder_type in text|image|audio; transformer_layer.py:203-209)."""
        if self.attn_ln is not None:
            raise NotImplementedError("scale_attn=True is not used by the 4B config (finetune_3B.yaml:128)")
        if self.training and (self.dropout_prob > 0 or self.drop_path_prob > 0):
            raise NotImplementedError("training-time dropout / drop-path: backward pass is not built yet")
        d, F_ = self.embed_dim, self.ffn_embed_dim
        M = B * S
        n = self._norm_pack()
        a = self.self_attn.pack()
        f = self._ffn_pack(modality)
        dev = x.device
        h = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
        K.layernorm(x, n["ln1_w"], n["ln1_b"], h, eps=self.self_attn_layer_norm.eps)
        o = self.self_attn.attend(h, bias, key_pad, B, S)
        K.gemm(o, a["wo"], K.EPI_RESID_F32, x, bias=a["bo"], gamma=n.get("g1"), resid=x)
        K.layernorm(x, n["ln2_w"], n["ln2_b"], h, eps=self.final_layer_norm.eps)
        u = torch.empty(M, F_, dtype=torch.bfloat16, device=dev)
        K.gemm(h, f["w01"], K.EPI_GEGLU_BF16, u)
        if "ln_w" in f:
            K.layernorm(u, f["ln_w"], f["ln_b"], u, eps=1e-5)
        K.gemm(u, f["w2"], K.EPI_RESID_F32, x, bias=f["b2"], gamma=n.get("g2"), resid=x)
        return x
