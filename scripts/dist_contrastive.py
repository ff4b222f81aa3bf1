"""BASELINE.json configs[3] (head-only form): image-text contrastive step on synthetic unit-norm
embeddings, local batch b per rank (default 1024), NCCL all-gather of both modalities, InfoNCE forward + backward through
the sm_100a kernels.  Launch:  torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_contrastive.py [--b 1024]
Checks rank 0's loss / gradients against oracle/restated.py on the concatenated batch, then times the step
(max over ranks, CUDA events) and prints one JSON line."""
import argparse, json, math, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import restated as R
import synth
from one_peace_b200.criterions.image_text_retrieval_loss import gather_without_grad, itc_loss

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=1024)
ap.add_argument("--d", type=int, default=1536)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
b, d = args.b, args.d
a_all, t_all = synth.contrastive_pair(b * world, d, seed=123)
a = a_all[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
t = t_all[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
ls = torch.tensor(math.log(1 / 0.07), device=dev, requires_grad=True)

def step():
    ga = gather_without_grad(a) if world > 1 else a.detach()
    gt = gather_without_grad(t) if world > 1 else t.detach()
    loss, i2t, t2i = itc_loss(a, t, ga, gt, ls.exp(), rank, 0.0)
    loss.backward()
    return loss, i2t, t2i

loss, i2t, t2i = step()
torch.cuda.synchronize()
ok = True
if rank == 0:
    ao = a_all[:b].clone().requires_grad_(True); to = t_all[:b].clone().requires_grad_(True)
    lo = torch.tensor(math.log(1 / 0.07), requires_grad=True)
    want, wi, wt = R.itc_loss(ao, to, a_all, t_all, R.logit_scale_exp(lo), 0, 0.0)
    want.backward()
    rel = abs(loss.item() - want.item()) / abs(want.item())
    gcos = torch.nn.functional.cosine_similarity(a.grad.cpu().flatten(), ao.grad.flatten(), dim=0).item()
    ok = rel < 1e-3 and float(i2t) == float(wi) and float(t2i) == float(wt) and gcos > 0.9995
    print(f"[rank0] loss {loss.item():.6f} oracle {want.item():.6f} rel {rel:.2e} | i2t {float(i2t)}/{float(wi)} | grad cos {gcos:.6f} | ok={ok}", flush=True)
for p in (a, t, ls):
    p.grad = None
for _ in range(3):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    per = ms.item() / args.steps
    print(json.dumps({"metric": "contrastive_head_pairs_per_sec", "value": round(b * world / (per / 1e3), 1), "unit": "pairs/s",
                      "n_gpus": world, "ms_per_step": round(per, 4), "config": {"workload": "InfoNCE fwd+bwd, head only", "local_batch": b,
                      "global_batch": b * world, "d": d}, "parity_ok": ok,
                      "flops_per_rank_per_step": 3 * 2 * 2.0 * b * (b * world) * (3 * d) }), flush=True)
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
