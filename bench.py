#!/usr/bin/env python
"""bench.py — headline measurement for the ONE-PEACE hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm: sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  (reference arm: the CPU path on host cores)

Workload (BASELINE.json configs[1], the configuration the metric is quoted on at one GPU):
    ONE-PEACE 4B vision-branch forward — `extract_image_features` on 64 synthetic 224x224 images per GPU:
    hMLP stem -> 40 x (d=1536, h=24, ffn=6144) encoder layers -> CLS LayerNorm -> image_proj -> L2 norm.
    bf16 operands / fp32 accumulate / fp32 residual stream.  Random-init weights of that architecture,
    synthetic inputs (no network for checkpoints or datasets).
A "step" is one such forward over one batch.  N > 1 = N independent data-parallel replicas (inference has
no exchange step: "replicas only", weak scaling); value = all samples / max-over-ranks device time.

One JSON line on stdout (rank 0).  Keys follow the driver contract; `roofline` describes the dominant
kernel (the tcgen05 GEMM), `cpu_baseline` the oracle port timed on the host cores, `e2e` the same metric
through the public API with pinned-host inputs (H2D + forward + D2H inside the timed region).

Extra blocks in the same line:
  contrastive         the path that HAS a collective (BASELINE.json configs[3]), at every N:
      head            InfoNCE fwd+bwd on synthetic unit-norm embeddings, local b = 1024 / rank: NCCL all-gather of both
                      modalities + fused loss / gradient kernels -> pairs/s, all-gather ms, parity vs the oracle on rank 0
      train_step      full image-text training step of the 4B text+image branches (encoder fwd/bwd, activations kept in HBM
                      when they fit + all-gather + InfoNCE + gradient reduce-scatter -> sharded fused Adam -> all-gather)
  gpu_eager_baseline  (N = 1) the reference's arithmetic (oracle/restated.py) in PyTorch eager bf16 on the same B200
                      (ATen / cuBLAS) for the config-2 forward AND the image-text training step (torch autograd + activation
                      checkpointing + fused torch AdamW): the "reference on the same box" bar
  hbm_kernels         (N = 1) CUDA-event GB/s of the HBM-bound kernels against MEASURED_PEAKS hbm_gbs
Skip them with --no-extras (the headline keys are unaffected).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 64
RES = 224
LAYERS, D, H, FFN = 40, 1536, 24, 6144
SEQ = (RES // 16) ** 2 + 1
METRIC = "encoder_samples_per_sec"
UNIT = "samples/s"


def _ncu_traffic_bytes():
    """DRAM bytes of the dominant GEMM launch (GeGLU epilogue, the `<2, 1, 1>` instantiation) from the committed ncu capture."""
    import csv
    path = os.path.join(ROOT, "profiles", "r02_ncu_layer_full.raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if "gemm_bf16_kernel<2, 1, 1>" in r[ik] or "gemm_bf16_kernel<(int)2, (int)1, (bool)1>" in r[ik]:
                return float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
    except Exception:
        pass
    return None


def _policy():
    from one_peace_b200 import autograd
    return autograd._POLICY


def flops_per_sample():
    """8d^2 + 6df + 4Sd per token per layer (+1.97 GFLOP hMLP stem)."""
    per_tok = 8 * D * D + 6 * D * FFN + 4 * SEQ * D
    return LAYERS * SEQ * per_tok + 1.97e9


def workload_config(n):
    return {
        "workload": "ONE-PEACE 4B vision-branch forward (extract_image_features): 64 x 3x224x224 per GPU, "
                    "40 layers d=1536 h=24 ffn=6144, S=197",
        "global_batch": BATCH * n, "seq_len": SEQ, "parallelism": f"dp{n} (independent replicas, no collective)",
        "l2": "no flush needed: one step streams 3.0 GB of bf16 weights + ~0.6 GB activations per layer group, "
              ">> 126 MB L2",
    }


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def build_model(device):
    import torch
    from one_peace_b200.one_peace import OnePeaceRetrievalConfig, OnePeaceRetrievalModel
    from one_peace_b200.unify_model_config import one_peace_4b_encoder_config
    cfg = OnePeaceRetrievalConfig()
    cfg.encoder = one_peace_4b_encoder_config(layers=LAYERS, embed_dim=D, ffn_embed_dim=FFN, attention_heads=H,
                                              patch_image_size=RES)
    torch.manual_seed(0)
    with torch.device(device):
        model = OnePeaceRetrievalModel(cfg, None, "image")
        # LayerScale at 1e-6 makes a fresh 4B model an identity map; use O(1) gammas and non-zero relpos
        # tables so the benchmark arithmetic is the same as with trained weights (timing is unaffected)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if "gamma_" in n:
                    p.fill_(0.1)
                elif "rel_pos_table" in n:
                    p.normal_(0, 0.1)
    model.eval()
    return model


def run_b200(args):
    import torch
    import torch.distributed as dist
    from one_peace_b200 import kernels as K
    from one_peace_b200.one_peace.hub_interface import OnePeaceHubInterface

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = build_model(dev)
    hub = OnePeaceHubInterface(model, device=dev)
    g = torch.Generator().manual_seed(1000 + rank)
    host_images = torch.randn(BATCH, 3, RES, RES, generator=g).pin_memory()
    dev_images = host_images.to(dev, non_blocking=True)
    host_out = torch.empty(BATCH, D, dtype=torch.float32).pin_memory()

    def step_core():
        with torch.no_grad():
            return model(src_images=dev_images, encoder_type="image")

    def step_e2e():
        # public API: host images in, host embeddings out
        return hub.extract_image_features(host_images, out=host_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        barrier()
        if sampler is not None:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = K.LAUNCHES
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler is not None else None
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, K.LAUNCHES - l0, clocks

    warmup = max(3, args.warmup)
    ms, launches, clocks = timed(step_core, args.steps, warmup, ClockSampler(local_rank) if rank == 0 else None)
    value = BATCH * world * args.steps / (ms / 1e3)
    ms_e2e, _, _ = timed(step_e2e, args.steps, warmup)
    e2e_value = BATCH * world * args.steps / (ms_e2e / 1e3)

    # ---- per-kernel device times (CUDA events on the launching stream) for the roofline block ----
    roof = None
    if rank == 0:
        recs = []
        state = {}

        def hook(kind, flops, shape):
            if kind == "gemm_begin":
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                state["e0"] = ev
            else:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                recs.append((state.pop("e0"), ev, flops, shape))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        K.PROFILE_HOOK = hook
        e0.record()
        step_core()
        e1.record()
        K.PROFILE_HOOK = None
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1)
        gemm_ms = sum(a.elapsed_time(b) for a, b, _, _ in recs)
        gemm_flops = sum(f for _, _, f, _ in recs)
        by_shape = {}
        for a, b, f, shp in recs:
            d = by_shape.setdefault(str(shp), [0.0, 0.0, 0])
            d[0] += a.elapsed_time(b); d[1] += f; d[2] += 1
        pk, pk_src = peaks()
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05, all four encoder GEMM shapes + stem/proj)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                "peak_source": pk_src + ", sustained figure (kernel timed inside a long step)",
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE GeGLU launch (the dominant shape, 12608x12288x1536)
                # from the `ncu --set full` capture summarised in profiles/r01_ncu_gemm_full_final.summary.txt; its
                # algorithmic bytes are A 38.7 + W 37.7 + out 154.9 = 231.4 MB: no wasted re-reads
                "traffic": _ncu_traffic_bytes(), "traffic_kernel": "gemm_bf16_kernel<2,GEGLU,TMA> (12608 x 12288 x 1536)",
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of that launch in profiles/r02_ncu_layer_full.raw.csv "
                                  "(ncu --set full; parsed at run time, not re-measured: ncu cannot run inside the bench); algorithmic "
                                  "bytes of the launch: A 38.7 + W 37.7 + out 154.9 = 231.4 MB",
                "launches": len(recs), "avg_launch_ms": round(gemm_ms / max(1, len(recs)), 4),
                "gemm_share_of_step": round(gemm_ms / step_ms, 4),
                "per_shape_tflops": {k: round(v[1] / (v[0] / 1e3) / 1e12, 1) for k, v in by_shape.items()},
                "whole_step_tflops": round(flops_per_sample() * BATCH / (ms / args.steps / 1e3) / 1e12, 1)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(steps=1, warmup=0, sample_images=8)

    extras = {}
    if not args.no_extras:
        del hub, model, dev_images
        torch.cuda.empty_cache()
        try:
            head = contrastive_head_block(dev, world, rank)
            train = contrastive_train_block(dev, world, rank)
            extras["contrastive"] = {"head": head, "train_step": train}
        except Exception as e:                      # the headline must survive a failure of an extra block
            extras["contrastive"] = {"error": repr(e)[:300]}
        if rank == 0 and world == 1:
            for key, fn in (("gpu_eager_baseline", eager_bf16_forward_baseline), ("hbm_kernels", hbm_kernels_block)):
                try:
                    extras[key] = fn(dev)
                except Exception as e:
                    extras[key] = {"error": repr(e)[:300]}
            if "forward" in extras.get("gpu_eager_baseline", {}):
                extras["gpu_eager_baseline"]["repo_over_eager_forward"] = round(value / extras["gpu_eager_baseline"]["forward"]["value"], 3)
                try:
                    et = eager_bf16_train_baseline(dev)
                    extras["gpu_eager_baseline"]["train_step"] = et
                    ts = extras.get("contrastive", {}).get("train_step", {})
                    if "value" in ts:
                        extras["gpu_eager_baseline"]["repo_over_eager_train_step"] = round(ts["value"] / et["value"], 3)
                except Exception as e:
                    extras["gpu_eager_baseline"]["train_step"] = {"error": repr(e)[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "ms_per_step": round(ms_e2e / args.steps, 3),
                    "h2d_bytes_per_step": host_images.numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4},
            "roofline": roof,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# extra blocks: contrastive head / train step (all N), eager-bf16 baseline and HBM-bound kernels (N = 1)
# ----------------------------------------------------------------------------------------------------
def _ev_ms(fn, iters, torch, pre=None):
    """Average device ms of fn() over `iters` launches (CUDA events on the current stream, one warm-up)."""
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def contrastive_head_block(dev, world, rank, b=1024, d=D, steps=20):
    """configs[3] head-only: all_gather of both (b, d) embedding matrices + InfoNCE fwd + bwd."""
    import math
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import restated as R
    import synth
    from one_peace_b200 import kernels as K
    from one_peace_b200.criterions.image_text_retrieval_loss import gather_without_grad, itc_loss
    a_all, t_all = synth.contrastive_pair(b * world, d, seed=123)
    a = a_all[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
    t = t_all[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
    ls = torch.tensor(math.log(1 / 0.07), device=dev, requires_grad=True)

    def gather():
        return (gather_without_grad(a), gather_without_grad(t)) if world > 1 else (a.detach(), t.detach())

    def step():
        ga, gt = gather()
        loss, i2t, t2i = itc_loss(a, t, ga, gt, ls.exp(), rank, 0.0)
        loss.backward()
        return loss, i2t, t2i
    loss, i2t, t2i = step()
    torch.cuda.synchronize()
    ok, rel = True, None
    if rank == 0:
        ao, to = a_all[:b].clone().requires_grad_(True), t_all[:b].clone().requires_grad_(True)
        lo = torch.tensor(math.log(1 / 0.07), requires_grad=True)
        torch.set_num_threads(min(os.cpu_count() or 1, 32))
        want, wi, wt = R.itc_loss(ao, to, a_all, t_all, R.logit_scale_exp(lo), 0, 0.0)
        want.backward()
        rel = abs(loss.item() - want.item()) / abs(want.item())
        gcos = torch.nn.functional.cosine_similarity(a.grad.cpu().flatten(), ao.grad.flatten(), dim=0).item()
        ok = bool(rel < 1e-3 and float(i2t) == float(wi) and float(t2i) == float(wt) and gcos > 0.9995)
    for q in (a, t, ls):
        q.grad = None
    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = K.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = (K.LAUNCHES - l0) // steps
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    ag = torch.tensor([_ev_ms(gather, 20, torch) if world > 1 else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ag, op=dist.ReduceOp.MAX)
    per = ms.item()
    n = b * world
    return {"metric": "contrastive_head_pairs_per_sec", "value": round(n / (per / 1e3), 1), "unit": "pairs/s",
            "ms_per_step": round(per, 4), "all_gather_ms": round(ag.item(), 4), "kernel_ms": round(per - ag.item(), 4),
            "local_batch": b, "global_batch": n, "d": d, "launches_per_step": int(launches), "parity_ok": ok,
            "loss_rel_vs_oracle": None if rel is None else float(f"{rel:.2e}"),
            "collective": "2 x all_gather_into_tensor (NCCL), forward only, rank-major (image_text_pretrain_loss.py:30-39)",
            "algorithmic_tflops_per_rank": round(3 * 2 * 2.0 * b * n * d / (per / 1e3) / 1e12, 1)}


def contrastive_train_block(dev, world, rank, b=64, text_len=32, steps=3, warmup=2):
    """configs[3] full-step form: the whole image-text training step of the 4B text + image branches."""
    import torch
    import torch.distributed as dist
    from one_peace_b200 import kernels as K
    from one_peace_b200.criterions import ImageTextRetrievalCriterion
    from one_peace_b200.one_peace import OnePeaceRetrievalConfig, OnePeaceRetrievalModel
    from one_peace_b200.one_peace.hub_interface import _Dictionary
    from one_peace_b200.optim.distributed_adam import DistributedAdam
    from one_peace_b200.unify_model_config import one_peace_4b_encoder_config
    cfg = OnePeaceRetrievalConfig()
    cfg.encoder = one_peace_4b_encoder_config(layers=LAYERS, embed_dim=D, ffn_embed_dim=FFN, attention_heads=H, patch_image_size=RES)
    torch.manual_seed(0)                      # identical initial weights on every rank
    with torch.device(dev):
        model = OnePeaceRetrievalModel(cfg, _Dictionary(50264), "vl")
        with torch.no_grad():
            for n, q in model.named_parameters():
                if "gamma_" in n:
                    q.fill_(0.1)
                elif "rel_pos_table" in n:
                    q.normal_(0, 0.1)
    model = model.to(torch.bfloat16)
    model.train()
    params = [q for q in model.parameters() if q.requires_grad]
    opt = DistributedAdam(params, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
    crit = ImageTextRetrievalCriterion(task=None, label_smoothing=0.0)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    sample = {"nsentences": b, "net_input": {
        "src_tokens": torch.randint(4, 50264, (b, text_len), device=dev, generator=g),
        "src_images": torch.randn(b, 3, RES, RES, device=dev, generator=g)}}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def step(timed=False):
        for q in params:
            q.grad = None
        if timed:
            ev[0].record()
        loss, _, _ = crit(model, sample)
        loss.backward()
        if timed:
            ev[1].record()
        opt.step(max_norm=3.0)               # clip_norm 3.0: pretrain_vl_3B.yaml
        if timed:
            ev[2].record()
        return loss.detach()
    losses = [round(step().item(), 4) for _ in range(warmup)]
    l0 = K.LAUNCHES
    step(timed=True)                              # eager step with the phase split (also the launch count)
    torch.cuda.synchronize()
    launches = K.LAUNCHES - l0
    eager_split = [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])]
    # the timed steps replay the criterion forward + backward as ONE CUDA graph (one_peace_b200/graphs.py; 4.9 k launches and
    # ~0.5 s of Python per eager step make the eager loop CPU-bound); the gradient exchange + optimizer stay eager
    graphed = None
    if os.environ.get("OPB_BENCH_GRAPH", "1") != "0":
        try:
            from one_peace_b200.graphs import GraphedTrainStep
            graphed = GraphedTrainStep(model, crit, sample, params, warmup=0)
        except Exception as e:
            graphed = None
            graph_error = repr(e)[:200]

    def run():
        if graphed is None:
            return step()
        loss, _, _ = graphed(sample)
        opt.step(max_norm=3.0)
        return loss.detach()
    run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        losses.append(round(run().item(), 4))
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps, eager_split[0], eager_split[1]], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n_params = sum(q.numel() for q in params)
    out = {"metric": "contrastive_train_step_pairs_per_sec", "value": round(b * world / t[0].item() * 1e3, 2), "unit": "pairs/s",
           "ms_per_step": round(t[0].item(), 2), "cuda_graph": graphed is not None,
           "eager_fwd_bwd_ms": round(t[1].item(), 2), "grad_exchange_adam_ms": round(t[2].item(), 2),
           "pairs_per_rank": b, "global_batch": b * world, "text_len": text_len, "params_b": round(n_params / 1e9, 3), "dtype": "bf16",
           "losses": losses, "finite": all(x == x for x in losses), "launches_per_step": int(launches),
           "max_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
           "collectives": "2 x all_gather (embeddings) + reduce_scatter(AVG) of the flat bf16 gradient + scalar all_reduce (norm) + "
                          "all_gather of the updated bf16 parameter shards (optim/distributed_adam.py); limiting one: the "
                          f"{round(n_params * 2 / 1e9, 2)} GB gradient reduce-scatter + parameter all-gather, not overlapped with backward",
           "activations": ("kept in HBM (no recompute: they fit in half of the free memory, autograd.keep_activations)"
                           if any(_policy().values()) else "recomputed per layer in the backward (checkpoint_wrapper policy)"),
           "includes": "text + image encoder fwd / bwd, InfoNCE, grad-norm clip, sharded fused Adam"}
    if graphed is None and os.environ.get("OPB_BENCH_GRAPH", "1") != "0":
        out["cuda_graph_error"] = graph_error
    del opt, model, params, graphed
    torch.cuda.empty_cache()
    return out


def eager_bf16_forward_baseline(dev, steps=5):
    """The reference's own arithmetic (oracle/restated.py = models/**/*.py restated, pinned by tests/golden) run as PyTorch eager
    bf16 on this GPU: bf16 weights / activations / residual stream, ATen + cuBLAS kernels, fp32 softmax as the reference does
    (multihead_attention.py:112).  Same workload as the headline: 64 x 224^2 images, 40 layers (4 distinct layers cycled)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import restated as R
    import synth
    distinct = 4
    sd = synth.make_state_dict(embed_dim=D, ffn=FFN, layers=distinct, heads=H, modalities=("image",), seed=0, gamma_range=(0.05, 0.15))
    sd = {k: (v.to(dev, torch.bfloat16) if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
    for i in range(distinct, LAYERS):
        for k in [k for k in sd if f"layers.{i % distinct}." in k]:
            sd[k.replace(f"layers.{i % distinct}.", f"layers.{i}.")] = sd[k]
    cfg = R.OracleConfig(embed_dim=D, ffn_embed_dim=FFN, layers=LAYERS, attention_heads=H)
    img = torch.randn(BATCH, 3, RES, RES, device=dev, generator=torch.Generator(device=dev).manual_seed(5)).to(torch.bfloat16)

    def fwd():
        with torch.no_grad():
            return R.extract_features(sd, cfg, "image", src_images=img)
    for _ in range(2):
        fwd()
    ms = _ev_ms(fwd, steps, torch)
    del sd
    torch.cuda.empty_cache()
    return {"forward": {"value": round(BATCH / (ms / 1e3), 2), "unit": UNIT, "ms_per_step": round(ms, 3)},
            "what": "oracle/restated.py (the reference's modules restated) in torch eager bf16 on the same GPU: ATen / cuBLAS, no "
                    "xformers / apex / flash-attn (the reference ships no Blackwell kernel)", "torch": torch.__version__}


def eager_bf16_train_baseline(dev, b=64, text_len=32, steps=3):
    """The reference's training arithmetic for the image-text contrastive step, as PyTorch eager bf16 on this GPU: oracle/restated.py
    text + image encoders (40 DISTINCT layers each branch's FFN, 2.73 B bf16 parameters drawn on the device), torch autograd with
    per-layer activation checkpointing (the recipes' checkpoint_activations), InfoNCE, clip_grad_norm_ and torch's fused Adam.
    No apex / xformers / flash-attn (not in the image; the reference ships no Blackwell kernel).  Same pairs, text length and
    parameter count as contrastive.train_step."""
    import torch
    from torch.utils.checkpoint import checkpoint
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import restated as R
    import synth
    sd1 = synth.make_state_dict(embed_dim=D, ffn=FFN, layers=1, heads=H, modalities=("text", "image"), seed=0, gamma_range=(0.05, 0.15))
    g = torch.Generator(device=dev).manual_seed(11)
    sd = {}
    for k, v in sd1.items():
        if "layers.0." in k:
            for i in range(LAYERS):
                t = v.to(dev, torch.bfloat16) if i == 0 else \
                    (torch.randn(v.shape, device=dev, generator=g) * float(v.float().std().clamp_min(1e-3)) + float(v.float().mean())).to(torch.bfloat16)
                sd[k.replace("layers.0.", f"layers.{i}.")] = t.requires_grad_(True)
        else:
            t = v.to(dev, torch.bfloat16) if v.is_floating_point() else v.to(dev)
            sd[k] = t.requires_grad_(True) if t.is_floating_point() else t
    cfg = R.OracleConfig(embed_dim=D, ffn_embed_dim=FFN, layers=LAYERS, attention_heads=H)
    params = [v for v in sd.values() if torch.is_tensor(v) and v.requires_grad]
    n_params = sum(q.numel() for q in params)
    opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05, fused=True)
    tok = torch.randint(4, 50264, (b, text_len), device=dev, generator=g)
    img = torch.randn(b, 3, RES, RES, device=dev, generator=g).to(torch.bfloat16)
    orig = R.encoder_layer

    def ckpt_layer(sd_, cfg_, x, bias, pad, modality, pfx):
        return checkpoint(lambda xx: orig(sd_, cfg_, xx, bias, pad, modality, pfx), x, use_reentrant=False)
    R.encoder_layer = ckpt_layer
    try:
        def step():
            opt.zero_grad(set_to_none=True)
            te = R.extract_features(sd, cfg, "text", src_tokens=tok)
            ie = R.extract_features(sd, cfg, "image", src_images=img)
            loss, _, _ = R.itc_loss(ie.float(), te.float(), ie.detach().float(), te.detach().float(),
                                    R.logit_scale_exp(sd["logit_scale"].float()), 0, 0.0)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 3.0, foreach=True)
            opt.step()
            return loss
        step()
        ms = _ev_ms(step, steps, torch)
    finally:
        R.encoder_layer = orig
    mem = round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)
    del opt, sd, params
    torch.cuda.empty_cache()
    return {"value": round(b / (ms / 1e3), 2), "unit": "pairs/s", "ms_per_step": round(ms, 2), "pairs": b, "text_len": text_len,
            "params_b": round(n_params / 1e9, 3), "max_mem_gb": mem,
            "what": "oracle/restated.py text + image branches, torch autograd + per-layer activation checkpointing, InfoNCE, "
                    "clip_grad_norm_, fused torch AdamW; eager bf16 (ATen / cuBLAS)"}


def hbm_kernels_block(dev):
    """Achieved GB/s of the HBM-bound kernels (CUDA events, buffers >> 126 MB L2) vs the measured copy bandwidth."""
    import ctypes
    import torch
    from one_peace_b200 import kernels as K
    from one_peace_b200.optim.adam import Adam
    pk, _ = peaks()
    hbm = pk["hbm_gbs"]
    out = []

    def rec(name, nbytes, ms, note):
        gbs = nbytes / (ms / 1e3) / 1e9
        out.append({"kernel": name, "algorithmic_mb": round(nbytes / 1e6, 1), "ms": round(ms, 4), "gbs": round(gbs, 1),
                    "frac": round(gbs / hbm, 3), "shape": note})
    bf, f32 = torch.bfloat16, torch.float32
    # fused Adam with fp32 master: 28 B / parameter; 1.511 B parameters = the 4B vision branch
    n = 1_511_000_000 // 8 * 8
    p = torch.nn.Parameter(torch.zeros(n, dtype=bf, device=dev))
    p.grad = torch.full((n,), 1e-3, dtype=bf, device=dev)
    opt = Adam([p], lr=1e-4, betas=(0.9, 0.98), weight_decay=0.05, master_weights=True)
    opt.step()
    rec("adam_multi_kernel", 28 * n, _ev_ms(opt.step, 5, torch), f"{n / 1e9:.3f} B bf16 params + fp32 master/m/v, 28 B/param")
    rec("grad_sumsq_kernel (+finalize)", 2 * n, _ev_ms(lambda: opt.grad_norm_and_scale(1.0, 3.0), 5, torch), f"{n / 1e9:.3f} B bf16 grads")
    del opt, p
    torch.cuda.empty_cache()
    rows, d = 4 * 12608, D
    x = torch.randn(rows, d, device=dev)
    w, b = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    y = torch.empty(rows, d, dtype=bf, device=dev)
    rec("layernorm_kernel fp32->bf16", rows * d * 6, _ev_ms(lambda: K.layernorm(x, w, b, y), 10, torch), f"[{rows}, {d}]")
    dy = torch.randn(rows, d, device=dev).to(bf)
    dx = torch.zeros(rows, d, device=dev)
    dg, db = torch.empty(d, device=dev), torch.empty(d, device=dev)
    rec("layernorm_bwd_kernel (+dgamma/dbeta)", rows * d * (4 + 2 + 4 + 4),
        _ev_ms(lambda: K.layernorm_bwd(x, dy, w, b, dx, accumulate=True, dgamma=dg, dbeta=db), 10, torch),
        f"x fp32 + dy bf16 in, dx fp32 read-modify-write, [{rows}, {d}]")
    u = torch.randn(rows // 2, 2 * FFN, device=dev).to(bf)
    uo = torch.empty(rows // 2, FFN, dtype=bf, device=dev)
    rec("geglu_fwd_kernel", (rows // 2) * FFN * 6, _ev_ms(lambda: K.geglu_fwd(u, uo), 10, torch), f"[{rows // 2}, 2 x {FFN}] -> [{rows // 2}, {FFN}] bf16")
    del u, uo
    Bt, T = 1024, 71
    tok = torch.randint(4, 50264, (Bt, T), device=dev)
    table = torch.randn(50264, d, device=dev).to(bf)
    pos, cls = torch.randn(514, d, device=dev), torch.randn(d, device=dev)
    rec("text_embed_kernel", Bt * (T + 1) * d * (2 + 4 + 4), _ev_ms(lambda: K.text_embed(tok, table, pos, cls), 10, torch),
        f"{Bt} x {T} tokens: bf16 table row + fp32 pos row read, fp32 row written")
    Ba, N = 16, 240000
    wav = torch.randn(Ba, N, device=dev)
    frames = (N - 10) // 5 + 1
    a0 = torch.empty(Ba * frames, 16, dtype=bf, device=dev)
    rec("audio_frame10_kernel", Ba * N * 4 + Ba * frames * 32, _ev_ms(lambda: K.audio_frame10(wav, frames, a0), 10, torch),
        f"{Ba} x 15 s waveform -> [{Ba * frames}, 16] bf16 frames")
    w0 = torch.randn(512, 16, device=dev).to(bf)
    y0 = torch.empty(Ba * frames, 512, dtype=bf, device=dev)
    rec("gemm_bf16_kernel (audio conv layer 0, K = 16)", Ba * frames * (32 + 1024), _ev_ms(lambda: K.gemm(a0, w0, K.EPI_STORE_BF16, y0), 10, torch),
        f"[{Ba * frames}, 16] x [512, 16]^T -> bf16 [{Ba * frames}, 512]: output-write bound")
    idx = torch.randperm(rows, device=dev)
    src = torch.empty(rows, 3 * d, dtype=bf, device=dev).normal_()
    dst = torch.empty_like(src)
    rec("row_gather_kernel (qkv modality-major -> batch-major)", rows * 3 * d * 4, _ev_ms(lambda: K.row_gather(src, idx, out=dst), 10, torch),
        f"[{rows}, {3 * d}] bf16 row permutation")
    tr = torch.empty(12608, 6144, dtype=bf, device=dev).normal_()
    rec("transpose_bf16_vec_kernel", 12608 * 6144 * 4, _ev_ms(lambda: K.transpose_bf16(tr), 10, torch), "[12608, 6144] bf16 (dW operand)")
    return {"peak_gbs": hbm, "kernels": out}


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (oracle/restated.py) on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, sample_images):
    """Times the CPU restatement of the reference path (fp32 torch CPU ops) on a bounded sample of the same
    workload: `sample_images` images through the full 40-layer vision branch.  /root/reference does not exist
    on the GPU box and the reference cannot be pip-installed (Python 3.12, missing omegaconf/hydra/...), so this
    is kind = "port" (the restatement is pinned to the reference by tests/test_oracle_golden.py).
    Thread count: the fastest of {all cores, 64, 32, 16} on a one-layer probe (torch's CPU GEMMs on a few hundred rows
    do not scale to 128 threads; taking the best setting keeps the baseline honest)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import restated as R
    import synth
    cores = os.cpu_count() or 1
    # distinct weights for 4 layers, cycled over the 40 (6 GB of fp32 weights would take minutes to draw)
    distinct = 4
    sd = synth.make_state_dict(embed_dim=D, ffn=FFN, layers=distinct, heads=H, modalities=("image",), seed=0)
    for i in range(distinct, LAYERS):
        for k in [k for k in sd if f"layers.{i % distinct}." in k]:
            sd[k.replace(f"layers.{i % distinct}.", f"layers.{i}.")] = sd[k]
    cfg = R.OracleConfig(embed_dim=D, ffn_embed_dim=FFN, layers=LAYERS, attention_heads=H)
    img = torch.randn(sample_images, 3, RES, RES, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        xp = torch.randn(sample_images, 197, D, generator=torch.Generator().manual_seed(6))
        padp = torch.zeros(sample_images, 197, dtype=torch.bool)
        best = (None, float("inf"))
        for th in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16)}, reverse=True):
            torch.set_num_threads(th)
            R.encoder_layer(sd, cfg, xp, None, padp, "image", "encoder_wrapper.fusion_model.layers.0.")
            dt = float("inf")
            for rep in range(3):            # best of three: a single probe made the two arms pick different counts (r01: 32 vs 16)
                t0 = time.perf_counter()
                R.encoder_layer(sd, cfg, xp, None, padp, "image", f"encoder_wrapper.fusion_model.layers.{1 + rep}.")
                dt = min(dt, time.perf_counter() - t0)
            if dt < best[1] * 0.97:         # ties go to the larger thread count tried first
                best = (th, dt)
        threads = best[0]
        torch.set_num_threads(threads)
        for _ in range(warmup):
            R.extract_features(sd, cfg, "image", src_images=img)
        t0 = time.perf_counter()
        for _ in range(steps):
            R.extract_features(sd, cfg, "image", src_images=img)
        dt = time.perf_counter() - t0
    v = sample_images * steps / dt
    return {"value": round(v, 3), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample_images} images x {steps} step(s) through the full 40-layer fp32 vision branch "
                      f"(oracle/restated.py, torch CPU ops, {threads} of {cores} host threads = fastest on a one-layer probe; "
                      f"layer weights cycled over {distinct} distinct sets)",
            "seconds": round(dt, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.gpus
    steps, warmup = max(1, min(args.steps, 2)), min(args.warmup, 1)
    r = cpu_reference(steps=steps, warmup=warmup, sample_images=8)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": n, "steps": steps,
            "warmup": warmup, "ms_per_step": round(1e3 * r["seconds"] / steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(n),
            "cpu_baseline": r, "gpu_launches": 0,
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the contrastive / eager-baseline / hbm_kernels blocks")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
