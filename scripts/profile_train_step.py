"""Kernel-time table of one image-text training step (config 4-ii) from torch.profiler (CUPTI): where the step goes.
usage: python scripts/profile_train_step.py [--layers 40] [--b 64] > profiles/rNN_train_step_kernels.txt"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from one_peace_b200.criterions import ImageTextRetrievalCriterion
from one_peace_b200.one_peace import OnePeaceRetrievalConfig, OnePeaceRetrievalModel
from one_peace_b200.one_peace.hub_interface import _Dictionary
from one_peace_b200.optim.adam import Adam
from one_peace_b200.unify_model_config import one_peace_4b_encoder_config

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=40)
ap.add_argument("--b", type=int, default=64)
args = ap.parse_args()
dev = torch.device("cuda")
cfg = OnePeaceRetrievalConfig()
cfg.encoder = one_peace_4b_encoder_config(layers=args.layers, embed_dim=1536, ffn_embed_dim=6144, attention_heads=24, patch_image_size=224)
torch.manual_seed(0)
with torch.device(dev):
    model = OnePeaceRetrievalModel(cfg, _Dictionary(50264), "vl")
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gamma_" in n:
                p.fill_(0.1)
            elif "rel_pos_table" in n:
                p.normal_(0, 0.1)
model = model.to(torch.bfloat16).train()
params = [p for p in model.parameters() if p.requires_grad]
opt = Adam(params, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
crit = ImageTextRetrievalCriterion(task=None, label_smoothing=0.0)
g = torch.Generator(device=dev).manual_seed(1)
sample = {"nsentences": args.b, "net_input": {"src_tokens": torch.randint(4, 50264, (args.b, 32), device=dev, generator=g),
                                              "src_images": torch.randn(args.b, 3, 224, 224, device=dev, generator=g)}}


def step():
    for p in params:
        p.grad = None
    loss, _, _ = crit(model, sample)
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"step (un-profiled): {e0.elapsed_time(e1):.1f} ms, layers={args.layers}, pairs={args.b}")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
