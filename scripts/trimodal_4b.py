"""BASELINE.json configs[2]: ONE-PEACE 4B tri-modal embedding on one B200 vs the CPU fp32 oracle — 8 images (224 x 224),
8 ragged token sequences (<= 71 tokens), 8 audio clips of 10 s (160000 samples -> 499 frames + CLS, two zero-padded with
their padding masks set); image-text and audio-text InfoNCE losses with logit_scale = ln(1 / 0.07).
Weights: seeded synthetic fp32 state dict at the 4B layer shape, 4 distinct layers cycled over the 40 (drawing 3.9 B
parameters takes minutes); the GPU model holds them in bf16.  Prints one JSON line."""
import json, math, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import restated as R
import synth
from one_peace_b200.criterions.image_text_retrieval_loss import itc_loss
from one_peace_b200.one_peace.hub_interface import from_pretrained

D, FFN, H, L, VOCAB, B = 1536, 6144, 24, 40, 4096, 8
if os.environ.get("OPB_TRIMODAL_TINY"):          # CPU dry-run of the oracle half (no GPU in the build container)
    D, FFN, H, L = 256, 1024, 4, 6
distinct = 4
sd = synth.make_state_dict(embed_dim=D, ffn=FFN, layers=distinct, heads=H, seed=2, vocab=VOCAB)
for i in range(distinct, L):
    for k in [k for k in sd if f"fusion_model.layers.{i % distinct}." in k]:
        sd[k.replace(f"fusion_model.layers.{i % distinct}.", f"fusion_model.layers.{i}.")] = sd[k]
g = torch.Generator().manual_seed(11)
tok = torch.randint(4, VOCAB, (B, 71), generator=g)
for i in range(B):
    tok[i, 71 - 7 * i:] = 1                                    # ragged: 71, 64, ... tokens, rest padding
img = torch.randn(B, 3, 224, 224, generator=g)
N = 160000
aud = torch.nn.functional.layer_norm(torch.randn(B, N, generator=g), (N,))
T = R.audio_frames(N, R.OracleConfig().feature_encoder_spec)
apm = torch.zeros(B, T + 1, dtype=torch.bool)
for b, keep in ((2, 0.6), (5, 0.35)):
    aud[b, int(N * keep):] = 0.0
    apm[b, 1 + int(T * keep):] = True

cfg = R.OracleConfig(embed_dim=D, ffn_embed_dim=FFN, layers=L, attention_heads=H)
torch.set_num_threads(min(os.cpu_count() or 1, 16))
t0 = time.perf_counter()
with torch.no_grad():
    wt = R.extract_features(sd, cfg, "text", src_tokens=tok)
    wi = R.extract_features(sd, cfg, "image", src_images=img)
    wa = R.extract_features(sd, cfg, "audio", src_audios=aud, audio_padding_masks=apm)
    scale = R.logit_scale_exp(sd["logit_scale"])
    w_itc, _, _ = R.itc_loss(wi, wt, wi, wt, scale, 0, 0.0)
    w_atc, _, _ = R.itc_loss(wa, wt, wa, wt, scale, 0, 0.0)
cpu_s = time.perf_counter() - t0
# second oracle pass with every floating-point parameter rounded to bf16 (what `dtype="bfloat16"` does to the model):
# separates weight quantisation — amplified by 40 random layers with O(1) LayerScale — from the kernels' own arithmetic
sdq = {k: (v.to(torch.bfloat16).float() if v.is_floating_point() else v) for k, v in sd.items()}
with torch.no_grad():
    qt = R.extract_features(sdq, cfg, "text", src_tokens=tok)
    qi = R.extract_features(sdq, cfg, "image", src_images=img)
    qa = R.extract_features(sdq, cfg, "audio", src_audios=aud, audio_padding_masks=apm)
    q_itc, _, _ = R.itc_loss(qi, qt, qi, qt, R.logit_scale_exp(sdq["logit_scale"]), 0, 0.0)
    q_atc, _, _ = R.itc_loss(qa, qt, qa, qt, R.logit_scale_exp(sdq["logit_scale"]), 0, 0.0)
if not torch.cuda.is_available():
    print("oracle half ok:", wt.shape, wi.shape, wa.shape, round(w_itc.item(), 4), round(w_atc.item(), 4), f"{cpu_s:.1f}s")
    sys.exit(0)

hub = from_pretrained(state_dict=sd, head_type="val", layers=L, embed_dim=D, ffn_embed_dim=FFN, attention_heads=H,
                      patch_image_size=224, device="cuda", dtype="bfloat16", vocab_size=VOCAB)
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / n
tc, ic, ac, pc = tok.cuda(), img.cuda(), aud.cuda(), apm.cuda()
gt, ms_t = timed(lambda: hub.extract_text_features(tc))
gi, ms_i = timed(lambda: hub.extract_image_features(ic))
ga, ms_a = timed(lambda: hub.extract_audio_features(ac, pc))
s = hub.model(return_logit_scale=True)
g_itc, _, _ = itc_loss(gi.float(), gt.float(), gi.float(), gt.float(), s.float(), 0, 0.0)
g_atc, _, _ = itc_loss(ga.float(), gt.float(), ga.float(), gt.float(), s.float(), 0, 0.0)
cos = lambda a, b: torch.nn.functional.cosine_similarity(a.float().cpu(), b).min().item()
def same_argmax(ga_, gb_, wa_, wb_, margin=2e-3):
    ws = wa_ @ wb_.t(); gs = ga_.float().cpu() @ gb_.float().cpu().t()
    top2 = ws.topk(2, dim=1).values
    dec = (top2[:, 0] - top2[:, 1]) > margin
    return bool((gs.argmax(1) == ws.argmax(1))[dec].all()), int(dec.sum())
i2t_ok, i2t_n = same_argmax(gi, gt, wi, wt)
a2t_ok, a2t_n = same_argmax(ga, gt, wa, wt)
rel = lambda a, b: abs(a.item() - b.item()) / abs(b.item())
line = {"config": "ONE-PEACE 4B tri-modal embedding, 8 images + 8 texts + 8 x 10 s audio, 1 x B200, bf16 weights vs fp32 CPU oracle",
        "min_cosine": {"text": round(cos(gt, wt), 6), "image": round(cos(gi, wi), 6), "audio": round(cos(ga, wa), 6)},
        "argmax_identical_on_decided_rows": {"i2t": [i2t_ok, i2t_n], "a2t": [a2t_ok, a2t_n]},
        "itc_loss": [round(g_itc.item(), 6), round(w_itc.item(), 6), f"rel {rel(g_itc, w_itc):.2e}"],
        "atc_loss": [round(g_atc.item(), 6), round(w_atc.item(), 6), f"rel {rel(g_atc, w_atc):.2e}"],
        "vs_oracle_with_bf16_rounded_weights": {
            "min_cosine": {"text": round(cos(gt, qt), 6), "image": round(cos(gi, qi), 6), "audio": round(cos(ga, qa), 6)},
            "text_cosine_per_row": [round(x, 5) for x in torch.nn.functional.cosine_similarity(gt.float().cpu(), qt).tolist()],
            "itc_loss_rel": f"{rel(g_itc, q_itc):.2e}", "atc_loss_rel": f"{rel(g_atc, q_atc):.2e}",
            "oracle_fp32_vs_oracle_bf16_weights_min_cosine": {"text": round(torch.nn.functional.cosine_similarity(wt, qt).min().item(), 6),
                                                              "image": round(torch.nn.functional.cosine_similarity(wi, qi).min().item(), 6),
                                                              "audio": round(torch.nn.functional.cosine_similarity(wa, qa).min().item(), 6)}},
        "gpu_ms": {"text": round(ms_t, 2), "image": round(ms_i, 2), "audio": round(ms_a, 2)},
        "samples_per_sec": {"text": round(B / ms_t * 1e3, 1), "image": round(B / ms_i * 1e3, 1), "audio": round(B / ms_a * 1e3, 1)},
        "cpu_oracle_seconds": round(cpu_s, 1), "audio_tokens": T + 1}
print(json.dumps(line))
