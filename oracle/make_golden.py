"""TEST INFRASTRUCTURE.  Generates tests/golden/*.pt by executing the reference's OWN module files
(via oracle/ref_stub.py) on seeded synthetic inputs.  Run in the build container only:

    python oracle/make_golden.py

The fixtures are small (tiny config: 2 layers, d=256, ffn=1024, 4 heads — BASELINE.json configs[0])
and are committed; the GPU box never sees /root/reference.
"""
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_stub  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def tiny_inputs():
    g = torch.Generator().manual_seed(0)
    tok = torch.randint(4, 50264, (32, 16), generator=g)
    for i in range(32):
        k = i % 4
        if k:
            tok[i, -k:] = 1
    img = torch.randn(2, 3, 224, 224, generator=g)
    aud = torch.randn(2, 16000, generator=g)
    aud = torch.nn.functional.layer_norm(aud, (16000,))
    aud[1, 12000:] = 0.0
    apm = torch.zeros(2, 50, dtype=torch.bool)       # 49 frames + cls
    apm[1, 38:] = True
    return tok, img, aud, apm


def randomise(model, seed=1):
    """The reference init leaves gamma=1e-6 / relpos tables 0 / biases 0, which hides most of the
    arithmetic; fixtures use a seeded perturbation of those so every term is exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gamma_" in n:
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif "rel_pos_table" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            elif n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif "layer_norm.weight" in n or n.endswith("ln.weight") or ".2.1.weight" in n or n.endswith("ffn.2.weight") \
                    or n.endswith("embed_audios.2.weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    model = ref_stub.build_reference_retrieval(embed_dim=256, ffn=1024, layers=2, heads=4, head_type="val", seed=0,
                                               vocab=50264)
    randomise(model)
    tok, img, aud, apm = tiny_inputs()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    # the 50264 x 256 token table dominates the file size: keep only the rows the fixture touches
    used = torch.unique(tok)
    emb_key = "encoder_wrapper.text_adapter.embed_tokens.weight"
    sd[emb_key + ".rows"] = used
    sd[emb_key + ".values"] = sd[emb_key][used].clone()
    del sd[emb_key]
    for k in list(sd):
        if k.endswith("rp_bucket") or k.endswith("position_idx") or k.endswith("version"):
            del sd[k]        # deterministic buffers, rebuilt by the oracle / product code
    with torch.no_grad():
        text = model(src_tokens=tok, encoder_type="text")
        image = model(src_images=img, encoder_type="image")
        audio = model(src_audios=aud, audio_padding_masks=apm, encoder_type="audio")
        # intermediate: adapter outputs and per-layer hidden states for the text branch
        tx, tpad, tbias = model.encoder_wrapper.text_adapter(tok)
        ix, ipad, ibias = model.encoder_wrapper.image_adapter(img)
        ax, apad, abias = model.encoder_wrapper.audio_adapter(aud, apm)
    torch.save({
        "config": dict(embed_dim=256, ffn_embed_dim=1024, layers=2, attention_heads=4, text_bucket_size=256,
                       image_bucket_size=16, image_rel_bucket_size=14, audio_bucket_size=512),
        "state_dict": sd,
        "inputs": dict(src_tokens=tok, src_images=img.half(), src_audios=aud.half(), audio_padding_masks=apm),
        "outputs": dict(text=text, image=image, audio=audio),
        "adapter": dict(text_x=tx, text_pad=tpad, text_bias=tbias[0][0], image_x=ix.half(), image_bias=ibias[0][0].half(),
                        audio_x=ax.half(), audio_bias=abias[0][0].half()),
    }, os.path.join(OUT, "tiny_retrieval.pt"))

    # ---- contrastive head (criterion file executed as-is; single process: .data path) ----
    crit_mod = ref_stub.ref_module("one_peace.criterions.image_text_retrieval_loss")
    g = torch.Generator().manual_seed(2)
    cases = []
    for (b, d, eps) in [(16, 64, 0.0), (48, 256, 0.1), (128, 1536, 0.0)]:
        img_e = torch.nn.functional.normalize(torch.randn(b, d, generator=g), dim=1)
        txt_e = torch.nn.functional.normalize(img_e + 0.5 * torch.nn.functional.normalize(torch.randn(b, d, generator=g), dim=1), dim=1)
        img_e.requires_grad_(True); txt_e.requires_grad_(True)
        ls = torch.tensor(math.log(1 / 0.07), requires_grad=True)
        crit = crit_mod.ImageTextRetrievalCriterion(task=None, label_smoothing=eps)
        scale = ls.exp()
        loss, i2t, t2i = crit.compute_itc_loss(img_e, txt_e, img_e.data, txt_e.data, scale)
        loss.backward()
        cases.append(dict(b=b, d=d, eps=eps, image=img_e.detach().clone(), text=txt_e.detach().clone(),
                          logit_scale=ls.detach().clone(), loss=loss.detach(), i2t_ncorrect=i2t, t2i_ncorrect=t2i,
                          grad_image=img_e.grad.clone(), grad_text=txt_e.grad.clone(), grad_logit_scale=ls.grad.clone()))
    torch.save(cases, os.path.join(OUT, "itc_loss.pt"))

    # ---- python Adam (optim/adam.py executed as-is) ----
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_adam", os.path.join(ref_stub.REF_PKG, "optim", "adam.py"))
    # adam.py imports fairseq.optim / omegaconf / dataclass helpers at module scope: provide shells
    import types
    sys.modules.setdefault("omegaconf", types.SimpleNamespace(II=lambda x: None, OmegaConf=object))
    sys.modules["fairseq.dataclass"].__dict__.setdefault("FairseqDataclass", object)
    for nm in ["one_peace.optim.adam_fused", "one_peace.optim.distributed_fused_adam", "one_peace.optim.base_optimizer"]:
        sys.modules.setdefault(nm, types.SimpleNamespace(FusedAdam=None, DistributedFusedAdam=None, BaseOptimizer=object))
    try:
        ref_adam = ref_stub.ref_module("one_peace.optim.adam")
        Adam = ref_adam.Adam
    except Exception as e:  # pragma: no cover
        raise RuntimeError(f"could not import the reference Adam: {e!r}")
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(1000, generator=g)
    grads = [torch.randn(1000, generator=g) * (0.1 + i) for i in range(3)]
    out = {}
    for tag, dt in [("fp32", torch.float32), ("bf16", torch.bfloat16)]:
        p = torch.nn.Parameter(p0.clone().to(dt))
        opt = Adam([p], lr=1e-2, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
        traj = []
        for gi in grads:
            p.grad = gi.clone().to(dt)
            opt.step()
            traj.append(p.detach().clone())
        st = opt.state[p]
        out[tag] = dict(p0=p0.clone().to(dt), grads=[gi.to(dt) for gi in grads], traj=traj, exp_avg=st["exp_avg"].clone(),
                        exp_avg_sq=st["exp_avg_sq"].clone(), lr=1e-2, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
    torch.save(out, os.path.join(OUT, "adam.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
