"""TEST INFRASTRUCTURE.  Deterministic synthetic weights and inputs shared by the golden-vector
generator (oracle/make_golden.py, build container), the CPU oracle tests and the GPU parity tests, so the
fixtures under tests/golden/ only need to hold OUTPUTS.  Everything is drawn from a seeded CPU
``torch.Generator`` in a fixed key order (same torch build here and on the GPU box).

``make_state_dict`` produces exactly the parameter names / shapes of the reference
``OnePeaceRetrievalModel`` (SURVEY.md §8a list; make_golden.py loads it with strict=True into the
reference model, which is the check that the list is right).
"""
import math

import torch
import torch.nn.functional as F


def _tn(g, shape, std=0.02):
    # truncated-normal-like: normal clipped at 2 sigma (shape of the reference's trunc_normal_ init)
    return (torch.randn(shape, generator=g) * std).clamp_(-2 * std, 2 * std)


def make_state_dict(embed_dim=256, ffn=1024, layers=2, heads=4, modalities=("text", "image", "audio"), seed=0,
                    vocab=50264, text_bucket=256, image_bucket=16, image_rel_bucket=14, audio_bucket=512,
                    conv_pos_depth=5, conv_pos_width=95, conv_pos_groups=16,
                    feature_spec=((512, 10, 5),) + ((512, 3, 2),) * 4 + ((512, 2, 2),) * 2, gamma_range=(0.5, 1.5)):
    g = torch.Generator().manual_seed(seed)
    d = embed_dim
    sd = {}

    def ln(prefix, n):
        sd[prefix + "weight"] = 1.0 + 0.2 * torch.randn(n, generator=g)
        sd[prefix + "bias"] = 0.1 * torch.randn(n, generator=g)

    def lin(prefix, out_f, in_f, bias=True, std=0.02):
        sd[prefix + "weight"] = _tn(g, (out_f, in_f), std)
        if bias:
            sd[prefix + "bias"] = 0.1 * torch.randn(out_f, generator=g)

    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    ew = "encoder_wrapper."
    if "text" in modalities:
        p = ew + "text_adapter."
        sd[p + "cls_embedding"] = _tn(g, (1, 1, d))
        emb = _tn(g, (vocab, d))
        emb[1] = 0
        sd[p + "embed_tokens.weight"] = emb
        sd[p + "embed_positions.weight"] = _tn(g, (514, d))
        sd[p + "rel_pos_table_list.0.weight"] = 0.5 * torch.randn(2 * text_bucket + 2, heads, generator=g)
    if "image" in modalities:
        p = ew + "image_adapter."
        sd[p + "cls_embedding"] = _tn(g, (1, 1, d))
        sd[p + "pos_embed"] = _tn(g, (image_bucket ** 2 + 1, d))
        c4 = d // 4
        sd[p + "embed_images.0.weight"] = torch.randn(c4, 3, 4, 4, generator=g) * (1.0 / math.sqrt(48))
        sd[p + "embed_images.0.bias"] = 0.1 * torch.randn(c4, generator=g)
        ln(p + "embed_images.1.layer_norm.", c4)
        sd[p + "embed_images.3.weight"] = torch.randn(c4, c4, 2, 2, generator=g) * (1.0 / math.sqrt(4 * c4))
        sd[p + "embed_images.3.bias"] = 0.1 * torch.randn(c4, generator=g)
        ln(p + "embed_images.4.layer_norm.", c4)
        sd[p + "embed_images.6.weight"] = torch.randn(d, c4, 2, 2, generator=g) * (0.5 / math.sqrt(4 * c4))
        sd[p + "embed_images.6.bias"] = 0.1 * torch.randn(d, generator=g)
        sd[p + "rel_pos_table_list.0.weight"] = 0.5 * torch.randn((2 * image_rel_bucket - 1) ** 2 + 3, heads, generator=g)
    if "audio" in modalities:
        p = ew + "audio_adapter."
        sd[p + "cls_embedding"] = _tn(g, (1, 1, d))
        sd[p + "cls_pos_embed"] = _tn(g, (1, 1, d))
        sd[p + "mask_embedding"] = _tn(g, (1, d))
        cin = 1
        for i, (c, k, s) in enumerate(feature_spec):
            sd[p + f"embed_audios.0.conv_layers.{i}.0.weight"] = torch.randn(c, cin, k, generator=g) * math.sqrt(2.0 / (cin * k))
            ln(p + f"embed_audios.0.conv_layers.{i}.2.1.", c)
            cin = c
        ln(p + "embed_audios.2.", cin)
        lin(p + "embed_audios.3.", d, cin, std=0.05)
        kpos = max(3, conv_pos_width // conv_pos_depth)
        for i in range(conv_pos_depth):
            sd[p + f"embed_positions.{i + 1}.0.weight"] = torch.randn(d, d // conv_pos_groups, kpos, generator=g) * \
                math.sqrt(1.0 / (d // conv_pos_groups * kpos))
            sd[p + f"embed_positions.{i + 1}.0.bias"] = 0.1 * torch.randn(d, generator=g)
        sd[p + "rel_pos_table_list.0.weight"] = 0.5 * torch.randn(2 * audio_bucket + 2, heads, generator=g)
    fm = ew + "fusion_model."
    for i in range(layers):
        p = fm + f"layers.{i}."
        lo, hi = gamma_range
        sd[p + "gamma_1"] = lo + (hi - lo) * torch.rand(d, generator=g)
        sd[p + "gamma_2"] = lo + (hi - lo) * torch.rand(d, generator=g)
        ln(p + "self_attn.ln.", d)
        lin(p + "self_attn.k_proj.", d, d, bias=False, std=0.05)
        lin(p + "self_attn.v_proj.", d, d, std=0.05)
        lin(p + "self_attn.q_proj.", d, d, std=0.05)
        lin(p + "self_attn.out_proj.", d, d, std=0.05)
        ln(p + "self_attn_layer_norm.", d)
        for m in modalities:
            q = p + f"{m}_ffn."
            lin(q + "0.wi_0.", ffn, d, bias=False, std=0.05)
            lin(q + "0.wi_1.", ffn, d, bias=False, std=0.05)
            ln(q + "2.", ffn)
            lin(q + "3.", d, ffn, std=0.03)
        ln(p + "final_layer_norm.", d)
    for m in modalities:
        ln(fm + f"{m}_layer_norm.", d)
    for m in modalities:
        lin(f"{m}_proj.", d, d, std=0.05)
    return sd


def tiny_inputs(seed=0, n_text=32, text_len=16, n_img=2, res=224, n_audio=2, audio_len=16000, vocab=50264):
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(4, vocab, (n_text, text_len), generator=g)
    for i in range(n_text):
        k = i % 4
        if k:
            tok[i, -k:] = 1
    img = torch.randn(n_img, 3, res, res, generator=g)
    aud = F.layer_norm(torch.randn(n_audio, audio_len, generator=g), (audio_len,))
    frames = audio_len
    for _, k, s in ((512, 10, 5),) + ((512, 3, 2),) * 4 + ((512, 2, 2),) * 2:
        frames = (frames - k) // s + 1
    apm = torch.zeros(n_audio, frames + 1, dtype=torch.bool)
    if n_audio > 1:
        aud[1, audio_len * 3 // 4:] = 0.0
        apm[1, 1 + frames * 3 // 4:] = True
    return tok, img, aud, apm


def contrastive_pair(b, d, seed, noise=0.5):
    """Unit-norm 'image' embeddings and correlated 'text' embeddings."""
    g = torch.Generator().manual_seed(seed)
    a = F.normalize(torch.randn(b, d, generator=g), dim=1)
    t = F.normalize(a + noise * F.normalize(torch.randn(b, d, generator=g), dim=1), dim=1)
    return a, t


def grad_summary(name, g):
    """Compact fingerprint of a gradient tensor: L2 norm, projection on a seeded random direction, first 256 values."""
    import zlib
    flat = g.detach().float().flatten()
    r = torch.randn(flat.numel(), generator=torch.Generator().manual_seed(zlib.crc32(name.encode())))
    return dict(shape=tuple(g.shape), norm=flat.norm().item(), proj=(flat * r).sum().item() / math.sqrt(flat.numel()),
                head=flat[:256].clone())


def retrieval_set(n_img, captions_per_image, d, seed, noise=0.8):
    """Unit-norm image embeddings, `captions_per_image` noisy text embeddings per image, and the id vectors the
    retrieval evaluation compares (text id = id of its image; metrics/recall.py:39-52)."""
    g = torch.Generator().manual_seed(seed)
    img = F.normalize(torch.randn(n_img, d, generator=g), dim=1)
    img_ids = torch.arange(100, 100 + n_img)
    txt = F.normalize(img.repeat_interleave(captions_per_image, 0) + noise * torch.randn(n_img * captions_per_image, d, generator=g) / math.sqrt(d) * 4, dim=1)
    txt_ids = img_ids.repeat_interleave(captions_per_image)
    perm = torch.randperm(txt.shape[0], generator=g)
    return img, txt[perm].contiguous(), img_ids, txt_ids[perm].contiguous()


# ----------------------------------------------------------------------------------------------------------------
# pretraining path (one_peace_pretrain.py + image_text_pretrain_loss.py): tiny encoder + decoder, masked sample
# ----------------------------------------------------------------------------------------------------------------
PRETRAIN_TINY = dict(embed_dim=256, ffn=1024, layers=2, heads=4, dec_dim=128, dec_ffn=256, dec_layers=2, dec_heads=2, res=64,
                     vocab=1000)


def make_pretrain_state_dict(embed_dim=256, ffn=1024, layers=2, heads=4, dec_dim=128, dec_ffn=256, dec_layers=2, dec_heads=2,
                             res=64, vocab=1000, seed=0, gamma_range=(0.5, 1.5)):
    """Parameter names / shapes of the reference ``OnePeacePretrainModel`` with text + image experts (pretrain_vl_3B.yaml
    structure at a tiny width): encoder as make_state_dict, decoder without LayerScale / relative-position bias / stems,
    projection heads, decoder_*_embed, *_mask_token, *_mask_head."""
    w = res // 16
    sd = make_state_dict(embed_dim=embed_dim, ffn=ffn, layers=layers, heads=heads, modalities=("text", "image"), seed=seed,
                         vocab=vocab, image_bucket=w, image_rel_bucket=w, gamma_range=gamma_range)
    dec = make_state_dict(embed_dim=dec_dim, ffn=dec_ffn, layers=dec_layers, heads=dec_heads, modalities=("text", "image"),
                          seed=seed + 1000, vocab=8, image_bucket=w, image_rel_bucket=w)
    for k, v in dec.items():
        if k.startswith("encoder_wrapper.fusion_model.") and ".gamma_" not in k:
            sd[k.replace("encoder_wrapper.", "decoder_wrapper.", 1)] = v
    for k in ("text_adapter.cls_embedding", "text_adapter.embed_positions.weight", "image_adapter.cls_embedding",
              "image_adapter.pos_embed"):
        sd["decoder_wrapper." + k] = dec["encoder_wrapper." + k]
    g = torch.Generator().manual_seed(seed + 2000)
    for m in ("text", "image"):
        sd[f"decoder_{m}_embed.weight"] = _tn(g, (dec_dim, embed_dim), 0.05)
        sd[f"decoder_{m}_embed.bias"] = 0.1 * torch.randn(dec_dim, generator=g)
        sd[f"{m}_mask_token"] = _tn(g, (1, dec_dim), 0.5)
        sd[f"{m}_mask_head.weight"] = _tn(g, (embed_dim, dec_dim), 0.08)
        sd[f"{m}_mask_head.bias"] = 0.1 * torch.randn(embed_dim, generator=g)
    return sd


def _preserve(mask_rows):
    """list of bool [S_i] mask vectors -> (mask (B,S) bool padded False, preserve_ids (B,K) int64 padded -1) as the collate
    function builds them (data/__init__.py:50-72)."""
    S = max(len(m) for m in mask_rows)
    ids = [(~m).nonzero(as_tuple=True)[0] for m in mask_rows]
    Kk = max(len(i) for i in ids)
    mask = torch.zeros(len(mask_rows), S, dtype=torch.bool)
    pres = torch.full((len(mask_rows), Kk), -1, dtype=torch.long)
    for b, (m, i) in enumerate(zip(mask_rows, ids)):
        mask[b, :len(m)] = m
        pres[b, :len(i)] = i
    return mask, pres


def pretrain_sample(seed=0, B=4, T=12, res=64, vocab=1000, text_mask_ratio=0.4, image_mask_ratio=0.75, vl_text_ratio=0.4,
                    vl_image_ratio=0.6875):
    """A collated image-text pretraining batch with the masking scheme of data/pretrain_data/image_text_pretrain_dataset.py
    :69-104 (token-level instead of whole-word masks): ragged captions ending in eos (2), pad id 1, no bos."""
    g = torch.Generator().manual_seed(seed)
    n_patch = (res // 16) ** 2
    tok = torch.ones(B, T, dtype=torch.long)
    tm, vtm, im, vim = [], [], [], []
    for b in range(B):
        n = T - 1 - (b % 3) * 2                                    # caption length without eos
        tok[b, :n] = torch.randint(4, vocab, (n,), generator=g)
        tok[b, n] = 2
        k = max(1, int(n * text_mask_ratio + 0.999))
        m = torch.zeros(n, dtype=torch.bool)
        m[torch.randperm(n, generator=g)[:k]] = True
        kv = int(n * vl_text_ratio)
        vm = torch.zeros(n, dtype=torch.bool)
        vm[torch.randn(n, generator=g).masked_fill(m, -float("inf")).argsort(descending=True)[:kv]] = True
        f = torch.zeros(1, dtype=torch.bool)
        tm.append(torch.cat([f, m, f]))
        vtm.append(torch.cat([f, vm, f]))
        mp = int(n_patch * image_mask_ratio)
        mi = torch.zeros(n_patch, dtype=torch.bool)
        mi[torch.randperm(n_patch, generator=g)[:mp]] = True
        vp = int(n_patch * vl_image_ratio)
        extra = torch.randn(n_patch, generator=g).masked_fill(~mi, -float("inf")).argsort(descending=True)[:vp - (n_patch - mp)]
        vmi = torch.zeros(n_patch, dtype=torch.bool)
        vmi[torch.cat([extra, (~mi).nonzero(as_tuple=True)[0]])] = True
        im.append(torch.cat([f, mi]))
        vim.append(torch.cat([f, vmi]))
    img = torch.randn(B, 3, res, res, generator=g)
    ni = {"src_tokens": tok, "src_images": img}
    for name, rows in (("text", tm), ("vl_text", vtm), ("image", im), ("vl_image", vim)):
        mask, pres = _preserve(rows)
        if name.endswith("text"):                                  # masks cover (B, T+1) positions
            full = torch.zeros(B, T + 1, dtype=torch.bool)
            full[:, :mask.shape[1]] = mask
            mask = full
        ni[f"{name}_mask_indices"], ni[f"{name}_preserve_ids"] = mask, pres
    return {"id": list(range(B)), "nsentences": B, "ntokens": B, "net_input": ni}


# ----------------------------------------------------------------------------------------------------------------
# audio-text pretraining path (one_peace_pretrain.py + audio_text_pretrain_loss.py; pretrain_al_3B.yaml structure)
# ----------------------------------------------------------------------------------------------------------------
PRETRAIN_AUDIO_TINY = dict(embed_dim=256, ffn=1024, layers=2, heads=4, dec_dim=128, dec_ffn=256, dec_layers=2, dec_heads=2,
                           vocab=1000)
AUDIO_SPEC = ((512, 10, 5),) + ((512, 3, 2),) * 4 + ((512, 2, 2),) * 2


def make_audio_pretrain_state_dict(embed_dim=256, ffn=1024, layers=2, heads=4, dec_dim=128, dec_ffn=256, dec_layers=2, dec_heads=2,
                                   vocab=1000, seed=0, gamma_range=(0.5, 1.5)):
    """Parameter names / shapes of the reference ``OnePeacePretrainModel`` with text + audio experts (pretrain_al_3B.yaml:90-170
    at a tiny width).  Decoder: no LayerScale, no relative-position bias; its audio adapter has no feature extractor and a
    learned Embedding(1026, d) position table (abs_pos_type 'fixed', models/adapter/audio.py:87-88)."""
    sd = make_state_dict(embed_dim=embed_dim, ffn=ffn, layers=layers, heads=heads, modalities=("text", "audio"), seed=seed,
                         vocab=vocab, gamma_range=gamma_range)
    dec = make_state_dict(embed_dim=dec_dim, ffn=dec_ffn, layers=dec_layers, heads=dec_heads, modalities=("text", "audio"),
                          seed=seed + 1000, vocab=8)
    for k, v in dec.items():
        if k.startswith("encoder_wrapper.fusion_model.") and ".gamma_" not in k:
            sd[k.replace("encoder_wrapper.", "decoder_wrapper.", 1)] = v
    for k in ("text_adapter.cls_embedding", "text_adapter.embed_positions.weight", "audio_adapter.cls_embedding",
              "audio_adapter.mask_embedding"):
        sd["decoder_wrapper." + k] = dec["encoder_wrapper." + k]
    g = torch.Generator().manual_seed(seed + 2000)
    sd["decoder_wrapper.audio_adapter.embed_positions.weight"] = _tn(g, (1026, dec_dim))
    for m in ("text", "audio"):
        sd[f"decoder_{m}_embed.weight"] = _tn(g, (dec_dim, embed_dim), 0.05)
        sd[f"decoder_{m}_embed.bias"] = 0.1 * torch.randn(dec_dim, generator=g)
        sd[f"{m}_mask_token"] = _tn(g, (1, dec_dim), 0.5)
        sd[f"{m}_mask_head.weight"] = _tn(g, (embed_dim, dec_dim), 0.08)
        sd[f"{m}_mask_head.bias"] = 0.1 * torch.randn(embed_dim, generator=g)
    return sd


def pretrain_audio_sample(seed=0, B=4, T=12, max_samples=16000, vocab=1000, audio_mask_ratio=0.55, al_text_ratio=0.4,
                          al_audio_ratio=0.45):
    """A collated audio-text pretraining batch shaped like data/pretrain_data/audio_text_pretrain_dataset.py:44-96 +
    data/__init__.py:59-81: ragged clips (zero-padded waveforms, frame padding mask), ragged captions ending in eos (2), pad 1;
    block masks replaced by seeded Bernoulli-style masks of fixed count."""
    g = torch.Generator().manual_seed(seed)
    tok = torch.ones(B, T, dtype=torch.long)
    wav = torch.zeros(B, max_samples)
    lens = [max_samples - (b % 3) * 2400 for b in range(B)]

    def frames(n):
        for _, k, s in AUDIO_SPEC:
            n = (n - k) // s + 1
        return n
    Tmax = frames(max_samples)
    apm = torch.zeros(B, Tmax + 1, dtype=torch.bool)
    vtm, am, vam = [], [], []
    f = torch.zeros(1, dtype=torch.bool)
    for b in range(B):
        n = T - 1 - (b % 3) * 2
        tok[b, :n] = torch.randint(4, vocab, (n,), generator=g)
        tok[b, n] = 2
        kv = max(1, int(n * al_text_ratio))
        vm = torch.zeros(n, dtype=torch.bool)
        vm[torch.randperm(n, generator=g)[:kv]] = True
        vtm.append(torch.cat([f, vm, f]))
        w = torch.randn(lens[b], generator=g)
        wav[b, :lens[b]] = (w - w.mean()) / w.std()
        Tb = frames(lens[b])
        apm[b, Tb + 1:] = True
        for ratio, dst in ((audio_mask_ratio, am), (al_audio_ratio, vam)):
            m = torch.zeros(Tb, dtype=torch.bool)
            m[torch.randperm(Tb, generator=g)[:int(Tb * ratio)]] = True
            dst.append(torch.cat([f, m]))
    ni = {"src_tokens": tok, "src_audios": wav, "audio_padding_masks": apm}
    for name, rows in (("al_text", vtm), ("audio", am), ("al_audio", vam)):
        mask, pres = _preserve(rows)
        width = T + 1 if name.endswith("text") else Tmax + 1
        full = torch.zeros(B, width, dtype=torch.bool)
        full[:, :mask.shape[1]] = mask
        ni[f"{name}_mask_indices"], ni[f"{name}_preserve_ids"] = full, pres
    return {"id": list(range(B)), "nsentences": B, "ntokens": B, "net_input": ni}
