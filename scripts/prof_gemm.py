"""Runs the four encoder-layer GEMMs (fused-LN epilogues, 4B shapes, M = 12608) a few times; used under ncu."""
import sys, torch
sys.path.insert(0, ".")
from one_peace_b200 import kernels as K
M, d, F = 12608, 1536, 6144
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
xb = torch.randn(M, d, device=dev, generator=g).bfloat16()
mu = torch.zeros(M, device=dev); rs = torch.ones(M, device=dev)
def w(n, k): return (torch.randn(n, k, device=dev, generator=g) * 0.03).bfloat16()
wqkv, wo, w01, w2 = w(3 * d, d), w(d, d), w(2 * F, d), w(d, F)
c3, b3, s3 = torch.randn(3 * d, device=dev), torch.randn(3 * d, device=dev), torch.ones(3 * d, device=dev)
c1, b1, g1 = torch.randn(d, device=dev), torch.randn(d, device=dev), torch.full((d,), 0.1, device=dev)
c2, b2 = torch.randn(2 * F, device=dev), torch.randn(2 * F, device=dev)
qkv = torch.empty(M, 3 * d, dtype=torch.bfloat16, device=dev)
o = torch.randn(M, d, device=dev, generator=g).bfloat16()
u = torch.empty(M, F, dtype=torch.bfloat16, device=dev)
x = torch.randn(M, d, device=dev, generator=g)
xb2 = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
part = torch.empty(48 * M * 2, device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    K.gemm_ln(xb, wqkv, K.EPI_STORE_BF16, qkv, ln_mu=mu, ln_rstd=rs, ln_colsum=c3, bias=b3, colscale=s3)
    K.gemm_ln(o, wo, K.EPI_RESID_F32, x, ln_mu=mu, ln_rstd=rs, ln_colsum=c1, bias=b1, gamma=g1, resid=x, stats_out=part, out_bf16=xb2)
    K.gemm_ln(xb, w01, K.EPI_GEGLU_BF16, u, ln_mu=mu, ln_rstd=rs, ln_colsum=c2, bias=b2, stats_out=part)
    K.gemm_ln(u, w2, K.EPI_RESID_F32, x, ln_mu=mu, ln_rstd=rs, ln_colsum=c1, bias=b1, gamma=g1, resid=x, stats_out=part, out_bf16=xb2)
torch.cuda.synchronize()
print("done")
