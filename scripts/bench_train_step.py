"""Forward + backward (+ fused Adam) of the 4B vision branch at the bench workload (64 x 224 x 224, M = 12608 rows,
40 layers): times the training path (one_peace_b200/autograd.py) with CUDA events and prints one JSON line.
Not the driver's bench contract (bench.py is) — a measurement of the backward kernels at full size.

    python scripts/bench_train_step.py [--steps 5] [--warmup 2] [--batch 64] [--layers 40] [--adam]
"""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from one_peace_b200 import kernels as K  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--layers", type=int, default=bench.LAYERS)
    ap.add_argument("--adam", action="store_true")
    args = ap.parse_args()
    bench.LAYERS = args.layers
    dev = torch.device("cuda:0")
    model = bench.build_model(dev).to(torch.bfloat16)
    model.train()
    g = torch.Generator(device=dev).manual_seed(0)
    img = torch.randn(args.batch, 3, bench.RES, bench.RES, device=dev, generator=g)
    target = torch.randn(args.batch, bench.D, device=dev, generator=g)
    opt = None
    if args.adam:
        from one_peace_b200.optim.adam import Adam
        opt = Adam([p for p in model.parameters() if p.requires_grad], lr=1e-5, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)

    ev = {"f0": [], "f1": []}

    def step(timed=False):
        for p in model.parameters():
            p.grad = None
        if timed:
            e0 = torch.cuda.Event(enable_timing=True); e0.record(); ev["f0"].append(e0)
        emb = model(src_images=img, encoder_type="image")
        if timed:
            e1 = torch.cuda.Event(enable_timing=True); e1.record(); ev["f1"].append(e1)
        loss = (emb.float() * target).sum()
        loss.backward()
        if opt is not None:
            opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = K.LAUNCHES
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(args.steps):
        loss = step(timed=True)
    e[1].record()
    torch.cuda.synchronize()
    launches = (K.LAUNCHES - launches0) // args.steps
    ms = e[0].elapsed_time(e[1]) / args.steps
    fwd_ms = sum(a.elapsed_time(b) for a, b in zip(ev["f0"], ev["f1"])) / args.steps
    flops_fwd = 38.69e12 * (args.batch / 64) * (args.layers / 40)
    # one step = forward + recompute + dX + dW GEMMs = 4 x forward GEMM work
    print(json.dumps({"metric": "train_step_samples_per_sec", "value": round(args.batch / ms * 1e3, 2), "unit": "samples/s",
                      "ms_per_step": round(ms, 3), "forward_ms": round(fwd_ms, 3),
                      "backward_incl_recompute_ms": round(ms - fwd_ms, 3), "adam": bool(args.adam), "batch": args.batch,
                      "layers": args.layers, "dtype": "bf16", "gemm_tflops_executed": round(4 * flops_fwd / ms / 1e9, 1),
                      "loss_finite": bool(torch.isfinite(loss).item()), "launches_per_step": launches,
                      "max_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))


if __name__ == "__main__":
    main()
