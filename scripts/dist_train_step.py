"""BASELINE.json configs[3], full-step form at a chosen local batch: image-text contrastive
TRAINING step of the 4B model on every rank —

    text encoder fwd (32 tokens) + image encoder fwd (224 x 224) -> NCCL all-gather of both embedding matrices ->
    fused InfoNCE loss / gradient -> hand-written encoder backward (activation recompute) ->
    gradient all-reduce (what fairseq's LegacyDDP does; here one flat NCCL all-reduce per dtype) -> fused Adam.

Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_train_step.py [--b 64]
Prints one JSON line (rank 0): pairs/s over all ranks, ms per step (max over ranks, CUDA events), loss trajectory."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from one_peace_b200 import kernels as K  # noqa: E402
from one_peace_b200.criterions import ImageTextRetrievalCriterion  # noqa: E402
from one_peace_b200.one_peace import OnePeaceRetrievalConfig, OnePeaceRetrievalModel  # noqa: E402
from one_peace_b200.one_peace.hub_interface import _Dictionary  # noqa: E402
from one_peace_b200.optim.adam import Adam  # noqa: E402
from one_peace_b200.unify_model_config import one_peace_4b_encoder_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=64, help="image-text pairs per rank")
ap.add_argument("--layers", type=int, default=40)
ap.add_argument("--text_len", type=int, default=32)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--vocab", type=int, default=50264)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lrk = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lrk)
dev = torch.device("cuda", lrk)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

cfg = OnePeaceRetrievalConfig()
cfg.encoder = one_peace_4b_encoder_config(layers=args.layers, embed_dim=1536, ffn_embed_dim=6144, attention_heads=24,
                                          patch_image_size=224)
torch.manual_seed(0)                      # identical initial weights on every rank
with torch.device(dev):
    model = OnePeaceRetrievalModel(cfg, _Dictionary(args.vocab), "vl")
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gamma_" in n:
                p.fill_(0.1)
            elif "rel_pos_table" in n:
                p.normal_(0, 0.1)
model = model.to(torch.bfloat16)
model.train()
params = [p for p in model.parameters() if p.requires_grad]
opt = Adam(params, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
crit = ImageTextRetrievalCriterion(task=None, label_smoothing=0.0)

g = torch.Generator(device=dev).manual_seed(1000 + rank)
sample = {"nsentences": args.b, "net_input": {
    "src_tokens": torch.randint(4, args.vocab, (args.b, args.text_len), device=dev, generator=g),
    "src_images": torch.randn(args.b, 3, 224, 224, device=dev, generator=g)}}


def allreduce_grads():
    """Average gradients over ranks (LegacyDistributedDataParallel's job in fairseq): flat buckets per dtype."""
    if world == 1:
        return
    by_dt = {}
    for p in params:
        if p.grad is not None:
            by_dt.setdefault(p.grad.dtype, []).append(p.grad)
    for dt, gs in by_dt.items():
        flat = torch.cat([x.reshape(-1) for x in gs])
        dist.all_reduce(flat)
        flat.div_(world)
        off = 0
        for x in gs:
            x.copy_(flat[off:off + x.numel()].view_as(x))
            off += x.numel()


def step():
    for p in params:
        p.grad = None
    loss, _, log = crit(model, sample)
    loss.backward()
    allreduce_grads()
    opt.step()
    return loss.detach(), log


losses = []
for _ in range(args.warmup):
    l, _ = step()
    losses.append(round(l.item(), 4))
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
l0 = K.LAUNCHES
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    l, log = step()
    losses.append(round(l.item(), 4))
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
n_params = sum(p.numel() for p in params)
if rank == 0:
    print(json.dumps({"metric": "contrastive_train_step_pairs_per_sec", "value": round(args.b * world / ms.item() * 1e3, 2),
                      "unit": "pairs/s", "n_gpus": world, "ms_per_step": round(ms.item(), 2), "pairs_per_rank": args.b,
                      "global_batch": args.b * world, "layers": args.layers, "text_len": args.text_len, "dtype": "bf16",
                      "params_b": round(n_params / 1e9, 3), "losses": losses, "finite": all(x == x for x in losses),
                      "i2t_ncorrect": float(log["i2t_ncorrect"]), "launches_per_step": (K.LAUNCHES - l0) // args.steps,
                      "max_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                      "includes": "text+image encoder fwd/bwd, all-gather, InfoNCE, grad all-reduce, fused Adam"}))
if world > 1:
    dist.destroy_process_group()
