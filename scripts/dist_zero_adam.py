"""ZeRO-1 sharded optimizer step (one_peace_b200/optim/distributed_adam.py) on N GPUs vs the un-sharded fused Adam on the
rank-averaged gradients.  Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_zero_adam.py
Every rank checks: identical parameters after 3 clipped steps (fp32 exactly the same arithmetic; bf16 with the fp32
master shard), global gradient norm == norm of the averaged gradients.  Prints ok=True on rank 0."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from one_peace_b200.optim.adam import Adam
from one_peace_b200.optim.distributed_adam import DistributedAdam

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lrk = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lrk)
dev = torch.device("cuda", lrk)
dist.init_process_group("nccl", device_id=dev)
shapes = [(1536, 1536), (1536,), (6144, 1536), (17,), (1, 1, 1536), (1000, 24), (3,)]
ok = True
for dt, tol in ((torch.float32, 2e-6), (torch.bfloat16, 1e-2)):
    g = torch.Generator(device=dev).manual_seed(7)
    base = [torch.randn(s, device=dev, generator=g) * 0.05 for s in shapes]
    pa = [torch.nn.Parameter(b.clone().to(dt)) for b in base]           # sharded
    pb = [torch.nn.Parameter(b.clone().to(dt)) for b in base]           # reference: plain fused Adam on averaged grads
    groups = lambda ps: [dict(params=ps[:3], weight_decay=0.05), dict(params=ps[3:], weight_decay=0.0, lr=2e-3)]
    oa = DistributedAdam(groups(pa), lr=1e-3, betas=(0.9, 0.98), eps=1e-8)
    ob = Adam(groups(pb), lr=1e-3, betas=(0.9, 0.98), eps=1e-8, master_weights=(dt != torch.float32))
    for step in range(3):
        gg = torch.Generator(device=dev).manual_seed(1000 * step + rank)
        grads = [(torch.randn(s, device=dev, generator=gg) * (3.0 if step == 1 else 0.3)).to(dt) for s in shapes]
        avg = []
        for gr in grads:
            a = gr.float().clone()
            dist.all_reduce(a)
            avg.append((a / world).to(dt))
        for p, gr in zip(pa, grads):
            p.grad = gr.clone()
        for p, gr in zip(pb, avg):
            p.grad = gr.clone()
        norm = oa.step(max_norm=1.0)
        ns = ob.grad_norm_and_scale(1.0, 1.0)
        ob.step(grad_scale=ns[1:2])
        want_norm = torch.sqrt(sum((a.float() ** 2).sum() for a in avg))
        rel = abs(norm.item() - want_norm.item()) / want_norm.item()
        ok = ok and rel < (1e-5 if dt == torch.float32 else 5e-3)
    worst = max(((a.detach().float() - b.detach().float()).abs().max() / (b.detach().float().abs().max() + 1e-12)).item()
                for a, b in zip(pa, pb))
    same = all(torch.equal(a.detach(), b) for a, b in zip(pa, [p.detach() for p in pa]))   # parameters are views of the flat buffer
    ok = ok and worst < tol and same
    if rank == 0:
        print(f"[{dt}] worst rel param diff vs un-sharded Adam {worst:.3e}, norm rel err {rel:.2e}, optimizer state "
              f"{oa.state_bytes_per_rank() / 1e6:.1f} MB / rank (un-sharded {sum(p.numel() for p in pa) * (12 if dt != torch.float32 else 8) / 1e6:.1f} MB)", flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"ok={bool(flag.item() == 1.0)}", flush=True)
dist.destroy_process_group()
