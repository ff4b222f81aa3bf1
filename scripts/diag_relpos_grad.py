"""Diagnostic: cosine of the image relative-position-table gradient (and the worst other tensor) of the tiny image-text InfoNCE
step vs the fp32 CPU oracle, for whatever OPB_* switches are set in the environment; with the eager-bf16 control."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import restated as R, synth
from one_peace_b200.criterions.image_text_retrieval_loss import itc_loss
from one_peace_b200.one_peace.hub_interface import from_pretrained
CFG = dict(embed_dim=256, ffn=1024, layers=2, heads=4)
mode = sys.argv[1] if len(sys.argv) > 1 else "itc"
sd = synth.make_state_dict(**CFG, modalities=("text", "image"), seed=4)
tok, img, _, _ = synth.tiny_inputs(seed=6, n_text=8, n_img=8, n_audio=1)
cfg = R.OracleConfig(embed_dim=CFG["embed_dim"], ffn_embed_dim=CFG["ffn"], layers=CFG["layers"], attention_heads=CFG["heads"])
g = torch.Generator().manual_seed(9)
target = torch.randn(8, CFG["embed_dim"], generator=g)


def oracle(sdx, dev, dt):
    sdg = {k: (v.detach().clone().to(dev, dt).requires_grad_(True) if v.is_floating_point() else v.to(dev)) for k, v in sdx.items()}
    te = R.extract_features(sdg, cfg, "text", src_tokens=tok.to(dev))
    ie = R.extract_features(sdg, cfg, "image", src_images=img.to(dev, dt))
    if mode == "itc":
        loss, _, _ = R.itc_loss(ie.float(), te.float(), ie.detach().float(), te.detach().float(), R.logit_scale_exp(sdg["logit_scale"].float()), 0, 0.0)
    else:
        loss = (ie.float() * target.to(dev)).sum()
    loss.backward()
    return {k: v.grad.float().cpu() for k, v in sdg.items() if v.is_floating_point() and v.grad is not None}, ie.detach().float().cpu()


want, ie_w = oracle(sd, "cpu", torch.float32)
eager, ie_e = oracle(sd, "cuda", torch.bfloat16)
hub = from_pretrained(state_dict=sd, head_type="vl", layers=2, embed_dim=256, ffn_embed_dim=1024, attention_heads=4, patch_image_size=224,
                      device="cuda", dtype="float32")
model = hub.model.train()
t = model(src_tokens=tok.cuda(), encoder_type="text")
i = model(src_images=img.cuda(), encoder_type="image")
if mode == "itc":
    loss, _, _ = itc_loss(i, t, i.detach(), t.detach(), model(return_logit_scale=True), 0, 0.0)
else:
    loss = (i.float() * target.cuda()).sum()
loss.backward()
cos = lambda a, b: torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()
key = "encoder_wrapper.image_adapter.rel_pos_table_list.0.weight"
got = {n: p.grad.float().cpu() for n, p in model.named_parameters() if p.grad is not None}
others = sorted((cos(got[n], want[n]), n) for n in got if n in want and n != key and want[n].abs().max() > 0 and "image" in n or "layers" in n and n in want and n in got)
print(f"[{mode}] switches: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("OPB_")))
print(f"   emb min cos: repo {torch.nn.functional.cosine_similarity(i.detach().float().cpu(), ie_w).min():.5f}  eager {torch.nn.functional.cosine_similarity(ie_e, ie_w).min():.5f}")
print(f"   relpos table: repo cos {cos(got[key], want[key]):.4f} |g|/|ref| {(got[key].norm() / want[key].norm()).item():.3f}   eager cos {cos(eager[key], want[key]):.4f} "
      f"|g|/|ref| {(eager[key].norm() / want[key].norm()).item():.3f}")
print("   worst other tensors (repo):", [(round(c, 4), n[-40:]) for c, n in others[:3]])
oe = sorted((cos(eager[n], want[n]), n) for n in eager if n in want and n != key and want[n].abs().max() > 0)
print("   worst other tensors (eager):", [(round(c, 4), n[-40:]) for c, n in oe[:3]])
# per-bucket view: which buckets carry the error
d = (got[key] - want[key]); w = want[key]
top = d.abs().sum(1).topk(5)
print("   buckets with the largest error:", [(int(ix), round(float(d[ix].abs().sum() / (w[ix].abs().sum() + 1e-12)), 3), round(float(w[ix].abs().sum() / w.abs().sum()), 4)) for ix in top.indices], "(bucket, rel err, share of |ref|)")
