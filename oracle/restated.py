"""TEST INFRASTRUCTURE — the CPU oracle.  Not product code; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.

A plain PyTorch fp32 restatement of the reference's algorithm on the hot path, written against a
``state_dict`` that uses the reference's own parameter names, so the same weights drive the reference
(in the build container, through oracle/ref_stub.py), this restatement (anywhere) and the CUDA path.

PINNING: tests/test_oracle_golden.py checks every function here against tests/golden/*.pt, which were
produced by executing the reference's unmodified module files (oracle/make_golden.py, run in the build
container where /root/reference exists).  The reference ships no tests / golden vectors of its own for
this path.

All file:line citations are relative to /root/reference/one_peace/.
"""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class OracleConfig:
    embed_dim: int = 1536
    ffn_embed_dim: int = 6144
    layers: int = 40
    attention_heads: int = 24
    text_bucket_size: int = 256
    image_bucket_size: int = 16        # pos_embed grid (run_scripts/finetune_3B.yaml: image_adapter.bucket_size)
    image_rel_bucket_size: int = 14    # = patch_image_size // 16 (one_peace_retrieval.py:128)
    audio_bucket_size: int = 512
    conv_pos_depth: int = 5
    conv_pos_width: int = 95
    conv_pos_groups: int = 16
    feature_encoder_spec: tuple = ((512, 10, 5),) + ((512, 3, 2),) * 4 + ((512, 2, 2),) * 2
    ln_eps: float = 1e-5
    pad_idx: int = 1


# ----------------------------------------------------------------------------------------------------
# relative-position buckets
# ----------------------------------------------------------------------------------------------------
def make_token_bucket_position(bucket_size, max_position=1024):
    """models/adapter/text.py:18-29 (audio twin: adapter/audio.py:20-32) + CLS overrides text.py:64-68."""
    ctx = torch.arange(max_position, dtype=torch.long)[:, None]
    mem = torch.arange(max_position, dtype=torch.long)[None, :]
    rel = ctx - mem
    sign = torch.sign(rel)
    mid = bucket_size // 2
    abs_pos = torch.where((rel < mid) & (rel > -mid), mid - 1, torch.abs(rel))
    log_pos = mid + torch.ceil(torch.log(abs_pos / mid) / math.log((max_position - 1) / mid) * (mid - 1)).long()
    bucket = torch.where(abs_pos.le(mid), rel, log_pos * sign).long() + bucket_size - 1
    n = 2 * bucket_size - 1
    bucket[0, :] = n
    bucket[:, 0] = n + 1
    bucket[0, 0] = n + 2
    return bucket


def make_image_bucket_position(w):
    """models/adapter/image.py:19-34 (BEiT-style 2-D index with 3 CLS ids)."""
    n = (2 * w - 1) * (2 * w - 1) + 3
    coords = torch.stack(torch.meshgrid([torch.arange(w), torch.arange(w)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rc = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rc[:, :, 0] += w - 1
    rc[:, :, 1] += w - 1
    rc[:, :, 0] *= 2 * w - 1
    idx = torch.zeros((w * w + 1,) * 2, dtype=rc.dtype)
    idx[1:, 1:] = rc.sum(-1)
    idx[0, 0:] = n - 3
    idx[0:, 0] = n - 2
    idx[0, 0] = n - 1
    return idx


def rel_pos_bias(table, bucket, seq_len):
    """adapter/text.py:84-91: table[bucket[:S,:S]] -> (S,S,H) -> (H,S,S); identical for every batch element."""
    return table[bucket[:seq_len, :seq_len]].permute(2, 0, 1)


# ----------------------------------------------------------------------------------------------------
# adapters: each returns (x (B,S,d), padding_mask (B,S) bool, bias (H,S,S))
# ----------------------------------------------------------------------------------------------------
def text_adapter(sd, cfg, src_tokens, prefix="encoder_wrapper.text_adapter."):
    """models/adapter/text.py:111-164 (no preserve_ids / mask-token branch: retrieval path)."""
    B, T = src_tokens.shape
    pad = torch.zeros(B, T + 1, dtype=torch.bool, device=src_tokens.device)
    pad[:, 1:] = src_tokens.eq(cfg.pad_idx)
    pos = sd[prefix + "embed_positions.weight"][: T + 1]
    emb = sd[prefix + "embed_tokens.weight"][src_tokens]
    cls = sd[prefix + "cls_embedding"].expand(B, -1, -1)
    x = torch.cat([cls, emb], dim=1) + pos[None]
    bucket = sd.get(prefix + "rp_bucket")
    if bucket is None:
        bucket = make_token_bucket_position(cfg.text_bucket_size)
    bias = rel_pos_bias(sd[prefix + "rel_pos_table_list.0.weight"], bucket, T + 1)
    return x, pad, bias


def image_adapter(sd, cfg, src_images, prefix="encoder_wrapper.image_adapter."):
    """models/adapter/image.py:206-260, hMLP stem :66-75, pos-embed resize :173-186."""
    B = src_images.size(0)
    w = src_images.size(2) // 16
    p = prefix + "embed_images."
    eps = cfg.ln_eps

    def ln2d(x, i):
        x = x.permute(0, 2, 3, 1)
        x = F.layer_norm(x, (x.size(-1),), sd[p + f"{i}.layer_norm.weight"], sd[p + f"{i}.layer_norm.bias"], eps)
        return x.permute(0, 3, 1, 2)

    x = F.conv2d(src_images, sd[p + "0.weight"], sd[p + "0.bias"], stride=4)
    x = F.gelu(ln2d(x, 1))
    x = F.conv2d(x, sd[p + "3.weight"], sd[p + "3.bias"], stride=2)
    x = F.gelu(ln2d(x, 4))
    x = F.conv2d(x, sd[p + "6.weight"], sd[p + "6.bias"], stride=2)
    x = x.flatten(2).transpose(1, 2)
    pos = image_pos_embed(sd[prefix + "pos_embed"], cfg.image_bucket_size, w)
    x = torch.cat([sd[prefix + "cls_embedding"].expand(B, -1, -1), x], dim=1) + pos[None]
    pad = torch.zeros(B, w * w + 1, dtype=torch.bool, device=src_images.device)
    bucket = sd.get(prefix + "rp_bucket")
    if bucket is None:
        bucket = make_image_bucket_position(cfg.image_rel_bucket_size)
    bias = sd[prefix + "rel_pos_table_list.0.weight"][bucket].permute(2, 0, 1)
    return x, pad, bias


def image_pos_embed(pos_embed, bucket_size, w):
    """models/adapter/image.py:173-186: bicubic resize (fp32, align_corners=False) unless w == bucket_size."""
    if w == bucket_size:
        return pos_embed
    cls_pos, old = pos_embed[:1], pos_embed[1:]
    old = old.reshape(1, bucket_size, bucket_size, -1).permute(0, 3, 1, 2).float()
    new = F.interpolate(old, size=(w, w), mode="bicubic").type_as(pos_embed)
    new = new.permute(0, 2, 3, 1).reshape(w * w, -1)
    return torch.cat([cls_pos, new], dim=0)


def audio_frames(n_samples, spec):
    """Frame count of the conv feature extractor (data/base_dataset.py:104-112; audio.py:254-311)."""
    L = n_samples
    for _, k, s in spec:
        L = (L - k) // s + 1
    return L


def audio_adapter(sd, cfg, src_audios, padding_mask, prefix="encoder_wrapper.audio_adapter."):
    """models/adapter/audio.py:150-210; feature extractor :254-311 (+ :46-55); conv-pos :57-80."""
    B = src_audios.size(0)
    eps = cfg.ln_eps
    x = src_audios.unsqueeze(1)
    for i, (dim, k, s) in enumerate(cfg.feature_encoder_spec):
        p = prefix + f"embed_audios.0.conv_layers.{i}."
        x = F.conv1d(x, sd[p + "0.weight"], None, stride=s)
        x = F.layer_norm(x.transpose(1, 2), (dim,), sd[p + "2.1.weight"], sd[p + "2.1.bias"], eps).transpose(1, 2)
        x = F.gelu(x)
    x = x.transpose(1, 2)
    x = F.layer_norm(x, (x.size(-1),), sd[prefix + "embed_audios.2.weight"], sd[prefix + "embed_audios.2.bias"], eps)
    feats = F.linear(x, sd[prefix + "embed_audios.3.weight"], sd[prefix + "embed_audios.3.bias"])
    # conv positional encoder on the un-normalised features (audio.py:194)
    kpos = max(3, cfg.conv_pos_width // cfg.conv_pos_depth)
    y = feats.transpose(1, 2)
    for i in range(cfg.conv_pos_depth):
        p = prefix + f"embed_positions.{i + 1}.0."
        y = F.conv1d(y, sd[p + "weight"], sd[p + "bias"], padding=kpos // 2, groups=cfg.conv_pos_groups)
        if kpos % 2 == 0:
            y = y[:, :, :-1]
        y = F.layer_norm(y.transpose(1, 2), (y.size(1),), None, None, eps).transpose(1, 2)
        y = F.gelu(y)
    pos = torch.cat([sd[prefix + "cls_pos_embed"].expand(B, -1, -1), y.transpose(1, 2)], dim=1)
    x = torch.cat([sd[prefix + "cls_embedding"].expand(B, -1, -1), feats], dim=1) + pos
    S = x.size(1)
    bucket = sd.get(prefix + "rp_bucket")
    if bucket is None:
        bucket = make_token_bucket_position(cfg.audio_bucket_size)
    bias = rel_pos_bias(sd[prefix + "rel_pos_table_list.0.weight"], bucket, S)
    return x, padding_mask, bias


# ----------------------------------------------------------------------------------------------------
# shared encoder
# ----------------------------------------------------------------------------------------------------
def attention(sd, cfg, x, bias, pad, p):
    """models/transformer/multihead_attention.py:103-126 (vanilla branch, magneto sub-LN, no c_attn)."""
    B, S, d = x.shape
    H = cfg.attention_heads
    hd = d // H
    q = F.linear(x, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"]) * hd ** -0.5
    k = F.linear(x, sd[p + "k_proj.weight"])
    v = F.linear(x, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    q, k, v = (t.view(B, S, H, hd).transpose(1, 2) for t in (q, k, v))
    a = q @ k.transpose(-1, -2)
    if bias is not None:
        a = a + bias[None]
    if pad is not None and pad.any():
        a = a.masked_fill(pad[:, None, None, :], float("-inf"))     # transformer_encoder.py:159-160
    o = (F.softmax(a, dim=-1, dtype=torch.float32).type_as(a) @ v).transpose(1, 2).reshape(B, S, d)
    o = F.layer_norm(o, (d,), sd[p + "ln.weight"], sd[p + "ln.bias"], cfg.ln_eps)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def geglu_ffn(sd, cfg, x, p):
    """models/transformer/transformer_layer.py:54-67,149-157: GeGLU -> LN(ffn) -> Linear."""
    u = F.gelu(F.linear(x, sd[p + "0.wi_0.weight"])) * F.linear(x, sd[p + "0.wi_1.weight"])
    u = F.layer_norm(u, (u.size(-1),), sd[p + "2.weight"], sd[p + "2.bias"], cfg.ln_eps)
    return F.linear(u, sd[p + "3.weight"], sd[p + "3.bias"])


def encoder_layer(sd, cfg, x, bias, pad, modality, p):
    """models/transformer/transformer_layer.py:165-228 (eval: dropout / drop-path off)."""
    d = x.size(-1)
    h = F.layer_norm(x, (d,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], cfg.ln_eps)
    x = x + sd[p + "gamma_1"] * attention(sd, cfg, h, bias, pad, p + "self_attn.")
    h = F.layer_norm(x, (d,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], cfg.ln_eps)
    x = x + sd[p + "gamma_2"] * geglu_ffn(sd, cfg, h, p + f"{modality}_ffn.")
    return x


def encoder(sd, cfg, x, pad, bias, modality, prefix="encoder_wrapper.fusion_model.", return_layer_inputs=False):
    """models/transformer/transformer_encoder.py:116-220, single-modality branches."""
    if pad.any():
        x = x * (1 - pad.unsqueeze(-1).type_as(x))          # :139-142
    states = []
    for i in range(cfg.layers):
        if return_layer_inputs:
            states.append(x)
        x = encoder_layer(sd, cfg, x, bias, pad, modality, prefix + f"layers.{i}.")
    d = x.size(-1)
    x = F.layer_norm(x, (d,), sd[prefix + f"{modality}_layer_norm.weight"], sd[prefix + f"{modality}_layer_norm.bias"],
                     cfg.ln_eps)
    return (x, states) if return_layer_inputs else x


def extract_features(sd, cfg, modality, src_tokens=None, src_images=None, src_audios=None, audio_padding_masks=None):
    """models/one_peace/one_peace_retrieval.py:86-123: adapter -> encoder -> CLS -> *_proj -> L2 normalise."""
    if modality == "text":
        x, pad, bias = text_adapter(sd, cfg, src_tokens)
    elif modality == "image":
        x, pad, bias = image_adapter(sd, cfg, src_images)
    elif modality == "audio":
        x, pad, bias = audio_adapter(sd, cfg, src_audios, audio_padding_masks)
    else:
        raise NotImplementedError(modality)
    x = encoder(sd, cfg, x, pad, bias, modality)
    cls = x[:, 0, :]
    return F.normalize(F.linear(cls, sd[f"{modality}_proj.weight"], sd[f"{modality}_proj.bias"]), dim=1)


def logit_scale_exp(logit_scale):
    """one_peace_retrieval.py:96-100: clamp to [0, ln 100] then exp."""
    return logit_scale.clamp(0, math.log(100)).exp()


# ----------------------------------------------------------------------------------------------------
# contrastive head
# ----------------------------------------------------------------------------------------------------
def label_smoothed_nll(lprobs, target, epsilon=0.0):
    """criterions/image_text_retrieval_loss.py:16-26 (mean over rows)."""
    nll = -lprobs.gather(dim=-1, index=target.unsqueeze(-1)).squeeze(-1)
    if epsilon != 0:
        smooth = -lprobs.sum(dim=-1)
        eps_i = epsilon / (lprobs.size(-1) - 1)
        loss = (1.0 - epsilon - eps_i) * nll + eps_i * smooth
    else:
        loss = nll
    return loss.mean()


def itc_loss(a_local, b_local, a_all, b_all, scale, rank=0, label_smoothing=0.0):
    """criterions/image_text_retrieval_loss.py:91-112.  a = image (or audio), b = text.
    Returns (loss, a2b_ncorrect, b2a_ncorrect).  *_all are detached gathers (:30-38)."""
    bsz = a_local.size(0)
    targets = torch.arange(bsz * rank, bsz * rank + bsz, device=a_local.device)
    sim_a2b = scale * a_local @ b_all.t()
    sim_b2a = scale * b_local @ a_all.t()
    la = F.log_softmax(sim_a2b, dim=-1, dtype=torch.float32).type_as(sim_a2b)
    lb = F.log_softmax(sim_b2a, dim=-1, dtype=torch.float32).type_as(sim_b2a)
    loss = (label_smoothed_nll(la, targets, label_smoothing) + label_smoothed_nll(lb, targets, label_smoothing)) / 2
    a_ok = (sim_a2b.argmax(dim=1) == targets).float().sum()
    b_ok = (sim_b2a.argmax(dim=1) == targets).float().sum()
    return loss, a_ok, b_ok


# ----------------------------------------------------------------------------------------------------
# optimizer
# ----------------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr, beta1, beta2, eps, weight_decay):
    """optim/adam.py:226-251 (python Adam: eps added to the un-bias-corrected sqrt(v), decoupled decay).
    fp32 tensors in, updated in place; `step` is the already-incremented step count."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    if weight_decay != 0:
        p.add_(p, alpha=-weight_decay * lr)
    p.addcdiv_(m, denom, value=-step_size)
    return p


def clip_coefficient(grads, max_norm, multiply_factor=1.0):
    """optim/fp16_optimizer_memory_efficent.py:96-116, bf16 branch: norm = factor * ||g||_2,
    coefficient = clamp(max_norm / (norm + 1e-6), max=1)."""
    total = torch.sqrt(sum((g.float() ** 2).sum() for g in grads))
    norm = multiply_factor * total
    coef = 1.0
    if max_norm > 0:
        coef = min(1.0, max_norm / (float(norm) + 1e-6))
    return float(norm), coef


# ----------------------------------------------------------------------------------------------------
# concatenated ('vl' / 'al') encoders and the DCL loss — SURVEY.md 8f "next" rows 1-2.  Restated and pinned now so
# that the CUDA side of the pretraining path has its checker; no product code uses them yet.
# ----------------------------------------------------------------------------------------------------
def encoder_multi(sd, cfg, parts, prefix="encoder_wrapper.fusion_model."):
    """models/transformer/transformer_encoder.py:116-232 for encoder_type 'vl' / 'al': `parts` = [(x, pad, bias,
    modality), ...] in sequence order (text first).  Sequences are concatenated, the per-modality relative-position
    biases sit on the diagonal blocks of one (H, S, S) bias (zero across modalities, :148-158), padded keys are -inf
    (:159-160); every layer runs shared attention and the FFN of each modality on its own rows
    (transformer_layer.py:203-219); each modality gets its own final LayerNorm (:207-220).  Returns the per-modality
    feature tensors."""
    x = torch.cat([p[0] for p in parts], dim=1)
    pad = torch.cat([p[1] for p in parts], dim=1)
    if pad.any():
        x = x * (1 - pad.unsqueeze(-1).type_as(x))
    H, S = cfg.attention_heads, x.size(1)
    bias = x.new_zeros(H, S, S)
    bounds, lo = [], 0
    for px, _, pb, _ in parts:
        hi = lo + px.size(1)
        if pb is not None:
            bias[:, lo:hi, lo:hi] += pb
        bounds.append((lo, hi))
        lo = hi
    d = x.size(-1)
    for i in range(cfg.layers):
        p = prefix + f"layers.{i}."
        h = F.layer_norm(x, (d,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], cfg.ln_eps)
        x = x + sd[p + "gamma_1"] * attention(sd, cfg, h, bias, pad, p + "self_attn.")
        h = F.layer_norm(x, (d,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], cfg.ln_eps)
        f = torch.cat([geglu_ffn(sd, cfg, h[:, a:b], p + f"{m}_ffn.") for (a, b), (_, _, _, m) in zip(bounds, parts)], dim=1)
        x = x + sd[p + "gamma_2"] * f
    outs = []
    for (a, b), (_, _, _, m) in zip(bounds, parts):
        outs.append(F.layer_norm(x[:, a:b], (d,), sd[prefix + f"{m}_layer_norm.weight"], sd[prefix + f"{m}_layer_norm.bias"],
                                 cfg.ln_eps))
    return outs


def dcl_loss(student_features, teacher_features, mask_indices, padding_masks=None, dcl_logit_scale=2.5, label_smoothing=0.1):
    """criterions/image_text_pretrain_loss.py:187-208: masked student tokens vs ALL (non-padded) teacher tokens of the local
    batch; CLS dropped; both sides L2-normalised in fp32; label-smoothed NLL, mean over the masked rows."""
    d = student_features.size(-1)
    stu = student_features[:, 1:, :].reshape(-1, d)
    tea = teacher_features.detach()[:, 1:, :].reshape(-1, d)
    mask = mask_indices[:, 1:].flatten()
    if padding_masks is not None:
        keep = torch.nonzero((~padding_masks).flatten(), as_tuple=False).flatten()
        stu, tea, mask = stu[keep], tea[keep], mask[keep]
    idx = torch.nonzero(mask, as_tuple=False).flatten()
    targets = torch.arange(stu.size(0))[idx]
    sim = dcl_logit_scale * F.normalize(stu[idx].float(), dim=1) @ F.normalize(tea.float(), dim=1).t()
    return label_smoothed_nll(F.log_softmax(sim, dim=-1, dtype=torch.float32), targets, label_smoothing)


# ----------------------------------------------------------------------------------------------------
# retrieval evaluation
# ----------------------------------------------------------------------------------------------------
def recall_eval(image_ids, image_logits, text_ids, text_logits):
    """metrics/recall.py:22-78 (single process): Recall@{1,5,10} both ways from the full similarity matrix."""
    sim_i2t = image_logits @ text_logits.t()
    out = {}
    for tag, scores, cand, own in (("txt", sim_i2t, text_ids, image_ids), ("img", sim_i2t.t(), image_ids, text_ids)):
        rank = scores.topk(k=min(10, scores.size(1)), dim=1).indices
        pred = cand[rank]
        rs = [100.0 * pred[:, :r].eq(own[:, None]).any(1).sum().item() / scores.size(0) for r in (1, 5, 10)]
        out[f"{tag}_r1"], out[f"{tag}_r5"], out[f"{tag}_r10"] = rs
        out[f"{tag}_r_mean"] = sum(rs) / 3
        out[f"predict_{tag}"] = pred
    out["r_mean"] = (out["txt_r_mean"] + out["img_r_mean"]) / 2
    return out


# ----------------------------------------------------------------------------------------------------
# pretraining model + criterion (models/one_peace/one_peace_pretrain.py:106-179, criterions/image_text_pretrain_loss.py:76-208)
# Pinned by tests/golden/pretrain_criterion.pt (the reference's own model + criterion executed on the same weights / batch).
# ----------------------------------------------------------------------------------------------------
def gather_bias(bias, position_ids):
    """adapter/text.py:96-101: (H,S,S) bias gathered on both axes by per-sample position ids (B,K) -> (B,H,K,K)."""
    B, Kk = position_ids.shape
    H, S = bias.shape[0], bias.shape[1]
    b = bias[None].expand(B, -1, -1, -1)
    b = b.gather(2, position_ids[:, None, :, None].expand(-1, H, -1, S))
    return b.gather(3, position_ids[:, None, None, :].expand(-1, H, Kk, -1))


def canvas(preserve_ids, preserve_embed, mask_token, seq_len):
    """adapter/text.py:135-142: mask token at every position, preserved rows scattered to their position ids."""
    B, Kk, d = preserve_embed.shape
    out = mask_token.repeat(B * seq_len, 1)
    right = torch.nonzero(preserve_ids.ne(-1).flatten(), as_tuple=False).flatten()
    left = (preserve_ids + (torch.arange(B, device=preserve_ids.device) * seq_len).unsqueeze(1)).view(-1)[right]
    out = out.index_put((left,), preserve_embed.reshape(-1, d)[right])
    return out.reshape(B, seq_len, d)


def text_adapter_general(sd, cfg, src_tokens, preserve_ids=None, preserve_embed=None, mask_token=None,
                         prefix="encoder_wrapper.text_adapter.", use_bias=True):
    """adapter/text.py:111-164 incl. the preserve_ids (:146-151) and mask-token (:135-142) branches."""
    B, T = src_tokens.shape
    S = T + 1
    pad = torch.zeros(B, S, dtype=torch.bool, device=src_tokens.device)
    pad[:, 1:] = src_tokens.eq(cfg.pad_idx)
    pos = sd[prefix + "embed_positions.weight"][:S][None].expand(B, -1, -1)
    bias = None
    if use_bias:
        bucket = sd.get(prefix + "rp_bucket")
        if bucket is None:
            bucket = make_token_bucket_position(cfg.text_bucket_size)
        bias = rel_pos_bias(sd[prefix + "rel_pos_table_list.0.weight"], bucket, S)
    if preserve_embed is not None:
        emb = canvas(preserve_ids, preserve_embed, mask_token, S)
    else:
        emb = torch.cat([sd[prefix + "cls_embedding"].expand(B, -1, -1), sd[prefix + "embed_tokens.weight"][src_tokens]], dim=1)
        if preserve_ids is not None:
            pad = preserve_ids.eq(-1)
            pid = preserve_ids.masked_fill(pad, preserve_ids.size(1) - 1)
            d = emb.size(-1)
            emb = emb.gather(1, pid[:, :, None].expand(-1, -1, d))
            pos = pos.gather(1, pid[:, :, None].expand(-1, -1, d))
            if bias is not None:
                bias = gather_bias(bias, pid)
    return emb + pos, pad, bias


def image_adapter_general(sd, cfg, src_images, preserve_ids=None, preserve_embed=None, mask_token=None,
                          prefix="encoder_wrapper.image_adapter.", use_bias=True):
    """adapter/image.py:206-260 incl. the preserve_ids (:241-246) and mask-token (:230-237) branches."""
    B = src_images.size(0)
    w = src_images.size(2) // 16
    S = w * w + 1
    pad = torch.zeros(B, S, dtype=torch.bool, device=src_images.device)
    pos = image_pos_embed(sd[prefix + "pos_embed"], cfg.image_bucket_size, w)[None].expand(B, -1, -1)
    bias = None
    if use_bias:
        bucket = sd.get(prefix + "rp_bucket")
        if bucket is None:
            bucket = make_image_bucket_position(cfg.image_rel_bucket_size)
        bias = sd[prefix + "rel_pos_table_list.0.weight"][bucket].permute(2, 0, 1)
    if preserve_embed is not None:
        emb = canvas(preserve_ids, preserve_embed, mask_token, S)
    else:
        x, _, _ = image_adapter(sd, cfg, src_images, prefix)          # includes + pos; undo to gather separately
        emb = x - pos
        if preserve_ids is not None:
            pad = preserve_ids.eq(-1)
            pid = preserve_ids.masked_fill(pad, preserve_ids.size(1) - 1)
            d = emb.size(-1)
            emb = emb.gather(1, pid[:, :, None].expand(-1, -1, d))
            pos = pos.gather(1, pid[:, :, None].expand(-1, -1, d))
            if bias is not None:
                bias = gather_bias(bias, pid)
    return emb + pos, pad, bias


def encoder_general(sd, cfg, parts, prefix="encoder_wrapper.fusion_model.", layer_scale=True):
    """models/transformer/transformer_encoder.py:116-232, every branch: parts = [(x (B,S_p,d), pad (B,S_p) bool, bias (H,S_p,S_p)
    | (B,H,S_p,S_p) | None, modality)].  Padded keys are excluded only through the bias (:159-160): without any bias they
    are attended.  `layer_scale=False`: no gamma_1 / gamma_2 (use_layer_scale off, the pretraining decoder)."""
    x = torch.cat([p[0] for p in parts], dim=1)
    pad = torch.cat([p[1] for p in parts], dim=1)
    if pad.any():
        x = x * (1 - pad.unsqueeze(-1).type_as(x))
    B, S, d = x.shape
    H = cfg.attention_heads
    bias = None
    bounds, lo = [], 0
    if any(p[2] is not None for p in parts):
        bias = x.new_zeros(B, H, S, S)
    for px, _, pb, _ in parts:
        hi = lo + px.size(1)
        if pb is not None:
            bias[:, :, lo:hi, lo:hi] += pb if pb.dim() == 4 else pb[None]
        bounds.append((lo, hi))
        lo = hi
    if bias is not None and pad.any():
        bias = bias.masked_fill(pad[:, None, None, :], float("-inf"))
    hd = d // H
    for i in range(cfg.layers):
        p = prefix + f"layers.{i}."
        h = F.layer_norm(x, (d,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], cfg.ln_eps)
        a = p + "self_attn."
        q = F.linear(h, sd[a + "q_proj.weight"], sd[a + "q_proj.bias"]) * hd ** -0.5
        k = F.linear(h, sd[a + "k_proj.weight"])
        v = F.linear(h, sd[a + "v_proj.weight"], sd[a + "v_proj.bias"])
        q, k, v = (t.view(B, S, H, hd).transpose(1, 2) for t in (q, k, v))
        att = q @ k.transpose(-1, -2)
        if bias is not None:
            att = att + bias
        o = (F.softmax(att, dim=-1, dtype=torch.float32).type_as(att) @ v).transpose(1, 2).reshape(B, S, d)
        o = F.layer_norm(o, (d,), sd[a + "ln.weight"], sd[a + "ln.bias"], cfg.ln_eps)
        o = F.linear(o, sd[a + "out_proj.weight"], sd[a + "out_proj.bias"])
        x = x + (sd[p + "gamma_1"] * o if layer_scale else o)
        h = F.layer_norm(x, (d,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], cfg.ln_eps)
        f = torch.cat([geglu_ffn(sd, cfg, h[:, a_:b_], p + f"{m}_ffn.") for (a_, b_), (_, _, _, m) in zip(bounds, parts)], dim=1)
        x = x + (sd[p + "gamma_2"] * f if layer_scale else f)
    outs = []
    for (a_, b_), (_, _, _, m) in zip(bounds, parts):
        outs.append(F.layer_norm(x[:, a_:b_], (d,), sd[prefix + f"{m}_layer_norm.weight"], sd[prefix + f"{m}_layer_norm.bias"],
                                 cfg.ln_eps))
    return outs


def audio_adapter_general(sd, cfg, src_audios, padding_mask, preserve_ids=None, preserve_embed=None, mask_token=None,
                          prefix="encoder_wrapper.audio_adapter.", use_bias=True):
    """adapter/audio.py:136-207 incl. the preserve_ids (:184-189; features gathered BEFORE the positional convolution, padded
    slots reading position K-1) and mask-token (:172-181; Embedding positions of the decoder variant) branches."""
    B, S = padding_mask.shape
    eps = cfg.ln_eps
    bias = None
    if use_bias:
        bucket = sd.get(prefix + "rp_bucket")
        if bucket is None:
            bucket = make_token_bucket_position(cfg.audio_bucket_size)
        bias = rel_pos_bias(sd[prefix + "rel_pos_table_list.0.weight"], bucket, S)
    if preserve_embed is not None:
        pos = sd[prefix + "embed_positions.weight"][:S][None].expand(B, -1, -1)
        return canvas(preserve_ids, preserve_embed, mask_token, S) + pos, padding_mask, bias
    x = src_audios.unsqueeze(1)
    for i, (dim, k, st) in enumerate(cfg.feature_encoder_spec):
        p = prefix + f"embed_audios.0.conv_layers.{i}."
        x = F.conv1d(x, sd[p + "0.weight"], None, stride=st)
        x = F.gelu(F.layer_norm(x.transpose(1, 2), (dim,), sd[p + "2.1.weight"], sd[p + "2.1.bias"], eps).transpose(1, 2))
    x = x.transpose(1, 2)
    x = F.layer_norm(x, (x.size(-1),), sd[prefix + "embed_audios.2.weight"], sd[prefix + "embed_audios.2.bias"], eps)
    feats = F.linear(x, sd[prefix + "embed_audios.3.weight"], sd[prefix + "embed_audios.3.bias"])
    if preserve_ids is not None:
        padding_mask = preserve_ids.eq(-1)
        pid = preserve_ids.masked_fill(padding_mask, preserve_ids.size(1) - 1)
        feats = feats.gather(1, pid[:, 1:, None].expand(-1, -1, feats.size(-1)) - 1)
        if bias is not None:
            bias = gather_bias(bias, pid)
    kpos = max(3, cfg.conv_pos_width // cfg.conv_pos_depth)
    y = feats.transpose(1, 2)
    for i in range(cfg.conv_pos_depth):
        p = prefix + f"embed_positions.{i + 1}.0."
        y = F.conv1d(y, sd[p + "weight"], sd[p + "bias"], padding=kpos // 2, groups=cfg.conv_pos_groups)
        if kpos % 2 == 0:
            y = y[:, :, :-1]
        y = F.gelu(F.layer_norm(y.transpose(1, 2), (y.size(1),), None, None, eps).transpose(1, 2))
    pos = torch.cat([sd[prefix + "cls_pos_embed"].expand(B, -1, -1), y.transpose(1, 2)], dim=1)
    return torch.cat([sd[prefix + "cls_embedding"].expand(B, -1, -1), feats], dim=1) + pos, padding_mask, bias


def pretrain_forward(sd, cfg, dec_cfg, src_tokens=None, text_preserve_ids=None, src_images=None, image_preserve_ids=None,
                     src_audios=None, audio_padding_masks=None, audio_preserve_ids=None, encoder_type=None):
    """models/one_peace/one_peace_pretrain.py:106-179.  cfg / dec_cfg: OracleConfig of the encoder / decoder."""
    parts = []
    if encoder_type in ("text", "vl", "al"):
        parts.append(text_adapter_general(sd, cfg, src_tokens, text_preserve_ids) + ("text",))
    if encoder_type in ("image", "vl"):
        parts.append(image_adapter_general(sd, cfg, src_images, image_preserve_ids) + ("image",))
    if encoder_type in ("audio", "al"):
        parts.append(audio_adapter_general(sd, cfg, src_audios, audio_padding_masks, audio_preserve_ids) + ("audio",))
    feats = dict(zip([p[3] for p in parts], encoder_general(sd, cfg, parts)))
    if text_preserve_ids is None and image_preserve_ids is None and audio_preserve_ids is None:
        if encoder_type in ("text", "image", "audio"):
            f = feats[encoder_type]
            logits = F.normalize(F.linear(f[:, 0, :], sd[f"{encoder_type}_proj.weight"], sd[f"{encoder_type}_proj.bias"]), dim=1)
            return logits, f
        return feats["text"], feats["image" if encoder_type == "vl" else "audio"]
    dparts = []
    if "text" in feats:
        emb = F.linear(feats["text"], sd["decoder_text_embed.weight"], sd["decoder_text_embed.bias"])
        dparts.append(text_adapter_general(sd, dec_cfg, src_tokens, text_preserve_ids, emb, sd["text_mask_token"],
                                           "decoder_wrapper.text_adapter.", use_bias=False) + ("text",))
    if "image" in feats:
        emb = F.linear(feats["image"], sd["decoder_image_embed.weight"], sd["decoder_image_embed.bias"])
        dparts.append(image_adapter_general(sd, dec_cfg, src_images, image_preserve_ids, emb, sd["image_mask_token"],
                                            "decoder_wrapper.image_adapter.", use_bias=False) + ("image",))
    if "audio" in feats:
        emb = F.linear(feats["audio"], sd["decoder_audio_embed.weight"], sd["decoder_audio_embed.bias"])
        dparts.append(audio_adapter_general(sd, dec_cfg, src_audios, audio_padding_masks, audio_preserve_ids, emb,
                                            sd["audio_mask_token"], "decoder_wrapper.audio_adapter.", use_bias=False) + ("audio",))
    dfeats = dict(zip([p[3] for p in dparts], encoder_general(sd, dec_cfg, dparts, "decoder_wrapper.fusion_model.",
                                                              layer_scale=False)))
    out = [None, None, None]
    for i, m in enumerate(("text", "image", "audio")):
        if m in dfeats:
            out[i] = F.linear(dfeats[m], sd[f"{m}_mask_head.weight"], sd[f"{m}_mask_head.bias"])
    return tuple(out)


def audio_text_pretrain_loss(sd, cfg, dec_cfg, net_input, alphas=(1.0, 0.5, 0.5), dcl_logit_scale=2.5, label_smoothing=0.0):
    """criterions/audio_text_pretrain_loss.py:73-158 (single process): the text tower is a frozen teacher (:94-95)."""
    ni = net_input
    tok, wav, apm = ni["src_tokens"], ni["src_audios"], ni["audio_padding_masks"]
    kw_a = dict(src_audios=wav, audio_padding_masks=apm)
    with torch.no_grad():
        text_logits, _ = pretrain_forward(sd, cfg, dec_cfg, src_tokens=tok, encoder_type="text")
    audio_logits, _ = pretrain_forward(sd, cfg, dec_cfg, encoder_type="audio", **kw_a)
    with torch.no_grad():
        teacher_al_text, teacher_al_audio = pretrain_forward(sd, cfg, dec_cfg, src_tokens=tok, encoder_type="al", **kw_a)
    _, _, student_audio = pretrain_forward(sd, cfg, dec_cfg, audio_preserve_ids=ni["audio_preserve_ids"], encoder_type="audio", **kw_a)
    sat, _, saa = pretrain_forward(sd, cfg, dec_cfg, src_tokens=tok, text_preserve_ids=ni["al_text_preserve_ids"],
                                   audio_preserve_ids=ni["al_audio_preserve_ids"], encoder_type="al", **kw_a)
    scale = logit_scale_exp(sd["logit_scale"])
    tpm, apad = tok.eq(cfg.pad_idx), apm[:, 1:]
    kw = dict(dcl_logit_scale=dcl_logit_scale, label_smoothing=label_smoothing)
    terms = {
        "dcl_audio_loss": dcl_loss(student_audio, teacher_al_audio, ni["audio_mask_indices"], apad, **kw),
        "dcl_al_text_loss": dcl_loss(sat, teacher_al_text, ni["al_text_mask_indices"], tpm, **kw),
        "dcl_al_audio_loss": dcl_loss(saa, teacher_al_audio, ni["al_audio_mask_indices"], apad, **kw),
    }
    atc, a2t, t2a = itc_loss(audio_logits, text_logits, audio_logits.detach(), text_logits.detach(), scale, 0, 0.0)
    terms["atc_loss"], terms["a2t_ncorrect"], terms["t2a_ncorrect"] = atc, a2t, t2a
    loss = atc + alphas[0] * terms["dcl_audio_loss"] + alphas[1] * terms["dcl_al_text_loss"] + alphas[2] * terms["dcl_al_audio_loss"]
    return loss, terms


def image_text_pretrain_loss(sd, cfg, dec_cfg, net_input, alphas=(0.5, 1.0, 0.5, 0.5), dcl_logit_scale=2.5, label_smoothing=0.0):
    """criterions/image_text_pretrain_loss.py:76-162 (single process).  Returns (loss, dict of the terms)."""
    ni = net_input
    tok, img = ni["src_tokens"], ni["src_images"]
    text_logits, teacher_text = pretrain_forward(sd, cfg, dec_cfg, src_tokens=tok, encoder_type="text")
    image_logits, teacher_image = pretrain_forward(sd, cfg, dec_cfg, src_images=img, encoder_type="image")
    with torch.no_grad():
        teacher_vl_text, teacher_vl_image = pretrain_forward(sd, cfg, dec_cfg, src_tokens=tok, src_images=img, encoder_type="vl")
    student_text, _, _ = pretrain_forward(sd, cfg, dec_cfg, src_tokens=tok, text_preserve_ids=ni["text_preserve_ids"],
                                          encoder_type="text")
    _, student_image, _ = pretrain_forward(sd, cfg, dec_cfg, src_images=img, image_preserve_ids=ni["image_preserve_ids"],
                                           encoder_type="image")
    svt, svi, _ = pretrain_forward(sd, cfg, dec_cfg, src_tokens=tok, text_preserve_ids=ni["vl_text_preserve_ids"], src_images=img,
                                   image_preserve_ids=ni["vl_image_preserve_ids"], encoder_type="vl")
    scale = logit_scale_exp(sd["logit_scale"])
    pm = tok.eq(cfg.pad_idx)
    kw = dict(dcl_logit_scale=dcl_logit_scale, label_smoothing=label_smoothing)
    terms = {
        "dcl_text_loss": dcl_loss(student_text, teacher_text, ni["text_mask_indices"], pm, **kw),
        "dcl_image_loss": dcl_loss(student_image, teacher_image, ni["image_mask_indices"], None, **kw),
        "dcl_vl_text_loss": dcl_loss(svt, teacher_vl_text, ni["vl_text_mask_indices"], pm, **kw),
        "dcl_vl_image_loss": dcl_loss(svi, teacher_vl_image, ni["vl_image_mask_indices"], None, **kw),
    }
    itc, i2t, t2i = itc_loss(image_logits, text_logits, image_logits.detach(), text_logits.detach(), scale, 0, 0.0)
    terms["itc_loss"], terms["i2t_ncorrect"], terms["t2i_ncorrect"] = itc, i2t, t2i
    loss = itc + alphas[0] * terms["dcl_text_loss"] + alphas[1] * terms["dcl_image_loss"] + \
        alphas[2] * terms["dcl_vl_text_loss"] + alphas[3] * terms["dcl_vl_image_loss"]
    return loss, terms
