"""One image-text PRETRAINING step at the 4B width (pretrain_vl_3B.yaml structure: 40-layer d=1536 encoder, 2-layer d=768 decoder):
`image_text_pretrain_loss` = ITC + 4 DCL terms, six model calls (teacher vl pass, contrastive text / image passes, three preserve_ids
student passes through encoder and decoder), backward, fused Adam.  Random-init bf16 weights, synthetic batch with the dataset's masking
ratios (oracle/synth.pretrain_sample).  usage: python scripts/bench_pretrain_step.py [B] [layers]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth
from one_peace_b200.criterions.image_text_pretrain_loss import ImageTextPretrainLossCriterion
from one_peace_b200.one_peace.hub_interface import from_pretrained
from one_peace_b200.optim.adam import Adam
from one_peace_b200 import kernels as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 40
RES, T = 256, 32
torch.manual_seed(0)
hub = from_pretrained(model_type="one_peace_pretrain", layers=L, embed_dim=1536, ffn_embed_dim=6144, attention_heads=24,
                      patch_image_size=RES, vocab_size=50264, device="cuda", dtype="bfloat16")
model = hub.model
with torch.no_grad():
    for n, p in model.named_parameters():
        if "gamma_" in n: p.fill_(0.1)
        elif "rel_pos_table" in n: p.normal_(0, 0.1)
model.train()
params = [p for p in model.parameters() if p.requires_grad]
opt = Adam(params, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
crit = ImageTextPretrainLossCriterion(None, label_smoothing=0.0)
sample = synth.pretrain_sample(seed=1, B=B, T=T, res=RES, vocab=50264)
sample = dict(sample, net_input={k: v.cuda() for k, v in sample["net_input"].items()})
sample["net_input"]["src_images"] = sample["net_input"]["src_images"].to(torch.bfloat16)


def step():
    for p in params:
        p.grad = None
    loss, _, log = crit(model, sample)
    loss.backward()
    opt.step()
    return loss.detach(), log


for _ in range(2):
    loss, log = step()
torch.cuda.synchronize()
l0 = K.LAUNCHES
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
n = 3
for _ in range(n):
    loss, log = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"image-text pretraining step, {B} pairs, {L}-layer 4B-width encoder + decoder, res {RES}, text {T}: {ms:.1f} ms/step "
      f"({B / ms * 1e3:.1f} pairs/s; host wall {1e3 * (time.perf_counter() - t0) / n:.0f} ms/step, eager), {(K.LAUNCHES - l0) // n} launches/step, "
      f"{sum(p.numel() for p in params) / 1e9:.2f} B parameters, peak {torch.cuda.max_memory_allocated() / 2**30:.1f} GB, loss {float(loss):.4f} "
      f"(itc {float(log['itc_loss']):.3f}, dcl {float(log['dcl_text_loss']):.3f} / {float(log['dcl_image_loss']):.3f} / "
      f"{float(log['dcl_vl_text_loss']):.3f} / {float(log['dcl_vl_image_loss']):.3f})")
if os.environ.get("OPB_PROFILE", "0") == "1":
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    ka = prof.key_averages()
    tot = sum(k.self_device_time_total for k in ka) / 1e3
    print(f"CUDA kernel time of one step (torch.profiler): {tot:.1f} ms; top kernels:")
    for k in sorted(ka, key=lambda k: -k.self_device_time_total)[:12]:
        print(f"   {k.self_device_time_total / 1e3:8.2f} ms  n={k.count:5d}  {k.key[:100]}")
