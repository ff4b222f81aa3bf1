"""Times the four encoder-layer GEMMs (fused-LN epilogues, 4B shapes, M = 12608) in layer order, repeatedly, with
CUDA events around each launch — in-situ conditions (clocks under sustained load, L2 state of a real step)."""
import os, sys, torch
sys.path.insert(0, ".")
from one_peace_b200 import kernels as K
M, d, F = 12608, 1536, 6144
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
xb = torch.randn(M, d, device=dev, generator=g).bfloat16()
mu = torch.zeros(M, device=dev); rs = torch.ones(M, device=dev)
def w(n, k): return (torch.randn(n, k, device=dev, generator=g) * 0.03).bfloat16()
NL = 6   # distinct weight sets so weights stream from HBM like in the real 40-layer loop
W = [(w(3 * d, d), w(d, d), w(2 * F, d), w(d, F)) for _ in range(NL)]
c3, b3, s3 = torch.randn(3 * d, device=dev), torch.randn(3 * d, device=dev), torch.ones(3 * d, device=dev)
c1, b1, g1 = torch.randn(d, device=dev), torch.randn(d, device=dev), torch.full((d,), 0.1, device=dev)
c2, b2 = torch.randn(2 * F, device=dev), torch.randn(2 * F, device=dev)
qkv = torch.empty(M, 3 * d, dtype=torch.bfloat16, device=dev)
o = torch.randn(M, d, device=dev, generator=g).bfloat16()
u = torch.empty(M, F, dtype=torch.bfloat16, device=dev)
x = torch.randn(M, d, device=dev, generator=g)
xb2 = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
part = torch.empty(96 * M * 2, device=dev)
tail_ws = torch.empty(256 * d, device=dev)
names = ["qkv", "out_proj", "geglu", "fc2"]
flops = [2.0 * M * 3 * d * d, 2.0 * M * d * d, 2.0 * M * 2 * F * d, 2.0 * M * d * F]
VARIANT = os.environ.get("OPB_EXP", "")
plain_out = [torch.empty(M, 3 * d, dtype=torch.bfloat16, device=dev), torch.empty(M, d, dtype=torch.bfloat16, device=dev),
             torch.empty(M, 2 * F, dtype=torch.bfloat16, device=dev), torch.empty(M, d, dtype=torch.bfloat16, device=dev)]
def layer(i, evs=None):
    wq, wo, w01, w2 = W[i % NL]
    if VARIANT == "plain":       # same mainloops, bare bf16-store epilogue (no LN / bias / GELU / residual): epilogue cost probe
        calls = [lambda: K.gemm_ln(xb, wq, K.EPI_STORE_BF16, plain_out[0]),
                 lambda: K.gemm_ln(o, wo, K.EPI_STORE_BF16, plain_out[1]),
                 lambda: K.gemm_ln(xb, w01, K.EPI_STORE_BF16, plain_out[2]),
                 lambda: K.gemm_ln(u, w2, K.EPI_STORE_BF16, plain_out[3])]
    else:
      calls = [lambda: K.gemm_ln(xb, wq, K.EPI_STORE_BF16, qkv, ln_mu=mu, ln_rstd=rs, ln_colsum=c3, bias=b3, colscale=s3),
             lambda: K.gemm_ln(o, wo, K.EPI_RESID_F32, x, ln_mu=mu, ln_rstd=rs, ln_colsum=c1, bias=b1, gamma=g1, resid=x, stats_out=part, out_bf16=xb2, workspace=tail_ws),
             lambda: K.gemm_ln(xb, w01, K.EPI_GEGLU_BF16, u, ln_mu=mu, ln_rstd=rs, ln_colsum=c2, bias=b2, stats_out=part),
             lambda: K.gemm_ln(u, w2, K.EPI_RESID_F32, x, ln_mu=mu, ln_rstd=rs, ln_colsum=c1, bias=b1, gamma=g1, resid=x, stats_out=part, out_bf16=xb2, workspace=tail_ws)]
    for k, f in enumerate(calls):
        if evs is not None:
            e0 = torch.cuda.Event(enable_timing=True); e0.record()
        f()
        if evs is not None:
            e1 = torch.cuda.Event(enable_timing=True); e1.record(); evs[k].append((e0, e1))
for i in range(40):
    layer(i)
torch.cuda.synchronize()
evs = [[], [], [], []]
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(120):
    layer(i, evs)
t1.record(); torch.cuda.synchronize()
tot = 0
out = []
for k in range(4):
    ms = sum(a.elapsed_time(b) for a, b in evs[k][40:]) / len(evs[k][40:])
    tot += ms
    out.append(f"{names[k]} {ms*1000:.0f}us {flops[k]/ms/1e9:.0f}TF")
# calibration: cuBLAS (torch.matmul, bf16, no epilogue work at all) on the same shapes, same in-situ loop
ref_out = [torch.empty(M, 3 * d, dtype=torch.bfloat16, device=dev), torch.empty(M, d, dtype=torch.bfloat16, device=dev),
           torch.empty(M, 2 * F, dtype=torch.bfloat16, device=dev), torch.empty(M, d, dtype=torch.bfloat16, device=dev)]
def ref_layer(i, evs=None):
    wq, wo, w01, w2 = W[i % NL]
    ops = [(xb, wq), (o, wo), (xb, w01), (u, w2)]
    for k, (a, b) in enumerate(ops):
        if evs is not None:
            e0 = torch.cuda.Event(enable_timing=True); e0.record()
        torch.matmul(a, b.t(), out=ref_out[k])
        if evs is not None:
            e1 = torch.cuda.Event(enable_timing=True); e1.record(); evs[k].append((e0, e1))
for i in range(20): ref_layer(i)
torch.cuda.synchronize()
revs = [[], [], [], []]
for i in range(80): ref_layer(i, revs)
torch.cuda.synchronize()
rtot = 0; rout = []
for k in range(4):
    ms = sum(a.elapsed_time(b) for a, b in revs[k][20:]) / len(revs[k][20:]); rtot += ms
    rout.append(f"{names[k]} {ms*1000:.0f}us {flops[k]/ms/1e9:.0f}TF")
print("cuBLAS (no epilogue): " + " | ".join(rout) + f" | layer {rtot*1000:.0f}us ({sum(flops)/rtot/1e9:.0f} TF)")
print(f"mode={os.environ.get('OPB_GEMM_TMA_EPILOGUE','-')} exp={VARIANT or '-'}: " + " | ".join(out) + f" | layer {tot*1000:.0f}us ({sum(flops)/tot/1e9:.0f} TF) wall/layer {t0.elapsed_time(t1)/120*1000:.0f}us")
