"""Drop-in for ``TransformerEncoderLayer`` / ``GeGLU`` (models/transformer/transformer_layer.py:54-228).

Parameter names match the reference (self_attn.*, self_attn_layer_norm, {text,image,audio}_ffn.{0.wi_0,
0.wi_1,2,3}, final_layer_norm, gamma_1, gamma_2).  One layer forward = 4 tcgen05 GEMMs + 1 attention
kernel + 4 LayerNorm kernels; the residual stream stays fp32 in HBM and is updated in place by the
out_proj / fc2 GEMM epilogues (gamma * (acc + bias) + residual — `fused_dropout_res`, :70-88, eval mode).
"""
import os

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

    # ------------------------------------------------------------------------------------------------
    # fused-LayerNorm path: the four LayerNorms of the layer never run as kernels.  Each GEMM consumes the
    # UN-normalised bf16 rows and applies  rstd * (acc - mu * colsum) + bias'  in its epilogue (gemm.h); the row
    # statistics come from the epilogue of the kernel that produced those rows.
    # ------------------------------------------------------------------------------------------------
    def fused_ln_supported(self):
        ffn_ok = all(isinstance(getattr(self, f"{m}_ffn")[2], nn.LayerNorm) for m in ("text", "image", "audio")
                     if hasattr(self, f"{m}_ffn"))
        return self.self_attn.ln is not None and ffn_ok and self.attn_ln is None and self.self_attn.c_attn is None

    @staticmethod
    def _fold(weights, ln, biases, interleave=False):
        """Folds LayerNorm `ln` into the Linear layers `weights` ([N_i, K] each, stacked along N; `interleave`: the GeGLU tile
        interleave of two weights) -> (bf16 W*diag(g) [sum N_i, K], colsum of the bf16 rows, bias' = W @ beta + b).
        One `opb_ln_fold` launch per source weight (csrc/pack.cu): the packs are rebuilt after every optimizer step."""
        dev = weights[0].device
        N, Kd = sum(w.shape[0] for w in weights), weights[0].shape[1]
        wg = torch.empty(N, Kd, dtype=torch.bfloat16, device=dev)
        colsum = torch.empty(N, dtype=torch.float32, device=dev)
        bias_out = torch.empty(N, dtype=torch.float32, device=dev)
        g, beta = f32(ln.weight), f32(ln.bias)
        off = 0
        for i, (w, b) in enumerate(zip(weights, biases)):
            wd = w.detach()
            if wd.dtype not in (torch.float32, torch.bfloat16):
                wd = wd.float()
            wd = wd.contiguous()
            bb = f32(b) if b is not None else None
            if interleave:
                K.ln_fold(wd, g, beta, bb, wg, colsum, bias_out, interleave=i + 1)
            else:
                n = w.shape[0]
                K.ln_fold(wd, g, beta, bb, wg[off:off + n], colsum[off:off + n], bias_out[off:off + n])
                off += n
        return wg, colsum, bias_out

    def _fused_attn_pack(self):
        cache = self._cache.setdefault("_fused_attn", PackCache())
        a = self.self_attn
        ps = [a.q_proj.weight, a.q_proj.bias, a.k_proj.weight, a.v_proj.weight, a.v_proj.bias, a.out_proj.weight,
              a.out_proj.bias, a.ln.weight, a.ln.bias, self.self_attn_layer_norm.weight, self.self_attn_layer_norm.bias] + \
             ([self.gamma_1] if self.gamma_1 is not None else [])

        def build():
            d = self.embed_dim
            dev = a.q_proj.weight.device
            w, c, dd = self._fold([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], self.self_attn_layer_norm,
                                  [a.q_proj.bias, None, a.v_proj.bias])
            qs = torch.ones(3 * d, device=dev)
            qs[:d] = a.scaling
            wo, co, do = self._fold([a.out_proj.weight], a.ln, [a.out_proj.bias])
            return dict(wqkv=w, cqkv=c, dqkv=dd, qscale=qs, wo=wo, co=co, do=do,
                        g1=f32(self.gamma_1) if self.gamma_1 is not None else None)
        return cache.get(ps, build)

    def _fused_ffn_pack(self, modality):
        cache = self._cache.setdefault("_fused_" + modality, PackCache())
        ffn = getattr(self, f"{modality}_ffn")
        ln2, lnf = self.final_layer_norm, ffn[2]
        ps = [ffn[0].wi_0.weight, ffn[0].wi_1.weight, ffn[3].weight, ffn[3].bias, lnf.weight, lnf.bias, ln2.weight,
              ln2.bias] + ([self.gamma_2] if self.gamma_2 is not None else [])

        def build():
            assert ffn[0].wi_0.weight.shape[0] % 128 == 0, "ffn_embed_dim must be a multiple of 128"
            w01, c01, d01 = self._fold([ffn[0].wi_0.weight, ffn[0].wi_1.weight], ln2, [None, None], interleave=True)
            w2, c2, d2 = self._fold([ffn[3].weight], lnf, [ffn[3].bias])
            return dict(w01=w01, c01=c01, d01=d01, w2=w2, c2=c2, d2=d2, lnf_eps=lnf.eps,
                        g2=f32(self.gamma_2) if self.gamma_2 is not None else None)
        return cache.get(ps, build)

    def forward_rows_fused(self, x, xb, ln1, ws, bias, key_pad, B, S, modality):
        """x fp32 [M,d] residual (in place), xb bf16 [M,d] copy of x, ln1 = LayerNorm-1 statistics of x: either
        dict(ln_mu=, ln_rstd=) (first layer) or dict(ln_partial=(records, parts, dim, eps)) (from the previous fc2).
        Returns the same for the layer output.  `ws` = workspace dict."""
        if self.training and (self.dropout_prob > 0 or self.drop_path_prob > 0):
            raise NotImplementedError("training-time dropout / drop-path: backward pass is not built yet")
        d, F_, H = self.embed_dim, self.ffn_embed_dim, self.self_attn.num_heads
        M = B * S
        a = self._fused_attn_pack()
        f = self._fused_ffn_pack(modality)
        n_t = (d + 255) // 256
        # LN1 -> QKV (+bias, q scale)
        K.gemm_ln(xb, a["wqkv"], K.EPI_STORE_BF16, ws["qkv"], ln_colsum=a["cqkv"], bias=a["dqkv"], colscale=a["qscale"],
                  workspace=ws["tail"], **ln1)          # workspace: split-K slabs when M < 256 (small batches)
        # attention; emits per-head partial statistics of its output rows (inner LN)
        self.self_attn.run_attention(ws["qkv"], bias, key_pad, B, S, out=ws["o"], ln_stats=ws["part_a"])
        # inner LN -> out_proj -> LayerScale + residual; emits x, xb and the partial statistics for LN2
        K.gemm_ln(ws["o"], a["wo"], K.EPI_RESID_F32, x, ln_partial=(ws["part_a"], H, d, self.self_attn.ln.eps),
                  ln_colsum=a["co"], bias=a["do"], gamma=a["g1"], resid=x, stats_out=ws["part_b"], out_bf16=xb,
                  workspace=ws["tail"])
        # LN2 -> GeGLU; emits u and the partial statistics for the FFN LayerNorm (96 records / row: reduced by a kernel)
        K.gemm_ln(xb, f["w01"], K.EPI_GEGLU_BF16, ws["u"], ln_partial=(ws["part_b"], n_t, d, self.final_layer_norm.eps),
                  ln_colsum=f["c01"], bias=f["d01"], stats_out=ws["part_c"])
        # FFN LN -> fc2 -> LayerScale + residual; emits x, xb and the partial statistics for the next layer's LN1.
        # OPB_FC2_INLINE_STATS=0 restores the separate ln_stats_finalize launch (96 records / row) for A/B runs.
        n_rec = 2 * ((2 * F_) // 256)                                                                     # 2 records / tile
        if os.environ.get("OPB_FC2_INLINE_STATS", "1") != "0":
            ln_ffn = dict(ln_partial=(ws["part_c"], n_rec, F_, f["lnf_eps"]))
        else:
            K.ln_stats_finalize(ws["part_c"], n_rec, M, F_, f["lnf_eps"], ws["mu2"], ws["rstd2"])
            ln_ffn = dict(ln_mu=ws["mu2"], ln_rstd=ws["rstd2"])
        K.gemm_ln(ws["u"], f["w2"], K.EPI_RESID_F32, x, ln_colsum=f["c2"], bias=f["d2"], gamma=f["g2"], resid=x,
                  stats_out=ws["part_d"], out_bf16=xb, workspace=ws["tail"], **ln_ffn)
        return dict(ln_partial=(ws["part_d"], n_t, d, self.self_attn_layer_norm.eps))

    @staticmethod
    def fused_workspace(M, d, F_, H, device):
        n_t = (d + 255) // 256
        e = lambda *s, dt=torch.bfloat16: torch.empty(*s, dtype=dt, device=device)
        f32 = torch.float32
        return dict(qkv=e(M, 3 * d), o=e(M, d), u=e(M, F_), xb=e(M, d), part_a=e(H * M * 2, dt=f32),
                    part_b=e(n_t * M * 2, dt=f32), part_c=e(2 * ((2 * F_) // 256) * M * 2, dt=f32), part_d=e(n_t * M * 2, dt=f32),
                    tail=e(16 * 256 * d, dt=torch.float32),      # split-K slabs of the residual GEMMs: one fp32 [256, d] slab per piece
                    mu=e(M, dt=torch.float32), rstd=e(M, dt=torch.float32), mu2=e(M, dt=torch.float32),
                    rstd2=e(M, dt=torch.float32))

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
            K.layernorm(u, f["ln_w"], f["ln_b"], u, eps=getattr(self, f"{modality}_ffn")[2].eps)
        K.gemm(u, f["w2"], K.EPI_RESID_F32, x, bias=f["b2"], gamma=n.get("g2"), resid=x)
        return x
