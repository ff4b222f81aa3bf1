"""Needs the extension built with `make -C one_peace_b200/csrc EXTRA=-DOPB_GEMM_TIMING` (touch gemm_tcgen05.cu first):
runs the two fp32-residual GEMMs of a 4B layer (out_proj K=1536, fc2 K=6144; M = 12608) and prints the per-phase cycle
accounting of their epilogue (opb_gemm_timing_dump)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from one_peace_b200 import kernels as K, _lib
lib = _lib.load()
M, d, F = 12608, 1536, 6144
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
mu = torch.zeros(M, device=dev); rs = torch.ones(M, device=dev)
def w(n, k): return (torch.randn(n, k, device=dev, generator=g) * 0.03).bfloat16()
wo, w2 = w(d, d), w(d, F)
c1, b1, g1 = torch.randn(d, device=dev), torch.randn(d, device=dev), torch.full((d,), 0.1, device=dev)
o = torch.randn(M, d, device=dev, generator=g).bfloat16()
u = torch.randn(M, F, device=dev, generator=g).bfloat16()
x = torch.randn(M, d, device=dev, generator=g)
xb = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
part = torch.empty(8 * M * 2, device=dev)
tail = torch.empty(256 * d, device=dev)
for name, a, wt in (("out_proj", o, wo), ("fc2", u, w2)):
    for _ in range(3):
        K.gemm_ln(a, wt, K.EPI_RESID_F32, x, ln_mu=mu, ln_rstd=rs, ln_colsum=c1, bias=b1, gamma=g1, resid=x, stats_out=part, out_bf16=xb, workspace=tail)
    lib.opb_gemm_timing_dump(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        K.gemm_ln(a, wt, K.EPI_RESID_F32, x, ln_mu=mu, ln_rstd=rs, ln_colsum=c1, bias=b1, gamma=g1, resid=x, stats_out=part, out_bf16=xb, workspace=tail)
    e1.record(); torch.cuda.synchronize()
    print(name, f"{e0.elapsed_time(e1) / 20 * 1000:.1f} us per launch", flush=True)
    lib.opb_gemm_timing_dump(1)
