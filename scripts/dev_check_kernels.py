"""Developer check run on the GPU box: every kernel vs a torch reference, prints max errors.
(The graded parity tests live in tests/; this script exists to get maximum signal out of one gpurun call.)"""
import sys, time, traceback
import torch
sys.path.insert(0, ".")
from one_peace_b200 import kernels as K

torch.manual_seed(0)
dev = "cuda"


def rel_err(a, b):
    a = a.float(); b = b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-9)).item()


def run(name, fn):
    try:
        t0 = time.time()
        fn()
        torch.cuda.synchronize()
        print(f"[ok  ] {name}  ({time.time()-t0:.2f}s)", flush=True)
    except Exception as e:
        print(f"[FAIL] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            print("  cuda context is dead:", e2, flush=True)
            sys.exit(3)


def gemm_case(M, N, K, cg, epi):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    ref = a.float() @ w.float().t()
    if epi == K_.EPI_STORE_BF16:
        cs = torch.rand(N, device=dev) + 0.5
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        K_.gemm(a, w, epi, out, bias=bias, colscale=cs, cta_group=cg)
        want = (ref + bias) * cs
    elif epi == K_.EPI_GELU_BF16:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        K_.gemm(a, w, epi, out, bias=bias, cta_group=cg)
        want = torch.nn.functional.gelu(ref + bias)
    elif epi == K_.EPI_GEGLU_BF16:
        out = torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16)
        K_.gemm(a, w, epi, out, cta_group=cg)
        r = ref.view(M, N // 256, 2, 128)
        want = (torch.nn.functional.gelu(r[:, :, 0]) * r[:, :, 1]).reshape(M, N // 2)
    elif epi == K_.EPI_RESID_F32:
        gamma = torch.randn(N, device=dev)
        resid = torch.randn(M, N, device=dev)
        out = resid.clone()
        K_.gemm(a, w, epi, out, bias=bias, gamma=gamma, resid=out, cta_group=cg)
        want = resid + gamma * (ref + bias)
    else:
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
        K_.gemm(a, w, epi, out, bias=bias, cta_group=cg)
        want = ref + bias
    torch.cuda.synchronize()
    e = rel_err(out, want)
    bad = (out.float() - want).abs() > 0.02 * want.abs().max()
    print(f"   gemm M={M} N={N} K={K} cg={cg} epi={epi}: rel_err={e:.3e} bad={int(bad.sum())}", flush=True)
    if e > 2e-2:
        idx = bad.nonzero()
        print("   first bad idx:", idx[:8].tolist(), " rows bad:", bad.any(1).sum().item(), "cols bad:", bad.any(0).sum().item())
        raise AssertionError(f"gemm mismatch {e}")


K_ = K


def attn_case(B, S, H, use_bias, use_pad):
    D = H * 64
    qkv = (torch.randn(B * S, 3 * D, device=dev) * 0.5).bfloat16()
    s_pad = (S + 7) // 8 * 8
    bias = None
    if use_bias:
        bias = torch.zeros(H, S, s_pad, device=dev)
        bias[:, :, :S] = torch.randn(H, S, S, device=dev)
    kp = None
    if use_pad:
        kp = torch.zeros(B, S, dtype=torch.uint8, device=dev)
        for b in range(B):
            n = (b * 3) % max(1, S // 2)
            if n:
                kp[b, S - n:] = 1
    out = K.attention(qkv, bias, kp, B, S, H)
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    sc = q @ k.transpose(-1, -2)
    if bias is not None:
        sc = sc + bias[None, :, :, :S]
    if kp is not None:
        sc = sc.masked_fill(kp.bool()[:, None, None, :], float("-inf"))
    p = sc.softmax(-1)
    want = (p @ v).permute(0, 2, 1, 3).reshape(B * S, D)
    torch.cuda.synchronize()
    e = rel_err(out, want)
    print(f"   attn B={B} S={S} H={H} bias={use_bias} pad={use_pad}: rel_err={e:.3e}", flush=True)
    assert e < 2e-2, e


def ln_case(rows, dim, in_dt, out_dt, affine, gelu, merge=0):
    x = torch.randn(rows, dim, device=dev) * 2 + 0.5
    x = x.to(in_dt)
    g = torch.randn(dim, device=dev) if affine else None
    b = torch.randn(dim, device=dev) if affine else None
    if merge:
        out = torch.zeros(rows // 4, dim * 4, device=dev, dtype=out_dt)
    else:
        out = torch.empty(rows, dim, device=dev, dtype=out_dt)
    K.layernorm(x, g, b, out, rows=rows, dim=dim, gelu=gelu, merge_grid_w=merge)
    want = torch.nn.functional.layer_norm(x.float(), (dim,), g, b, 1e-5)
    if gelu:
        want = torch.nn.functional.gelu(want)
    if merge:
        w = merge
        nb = rows // (w * w)
        want = want.view(nb, w // 2, 2, w // 2, 2, dim).permute(0, 1, 3, 2, 4, 5).reshape(rows // 4, 4 * dim)
    torch.cuda.synchronize()
    e = rel_err(out, want)
    print(f"   ln rows={rows} dim={dim} {in_dt}->{out_dt} affine={affine} gelu={gelu} merge={merge}: {e:.3e}", flush=True)
    assert e < 1.5e-2, e


def bench_gemm(M, N, K, cg, epi, iters=20):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    if epi == K_.EPI_GEGLU_BF16:
        out = torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16)
    elif epi == K_.EPI_RESID_F32:
        out = torch.zeros(M, N, device=dev, dtype=torch.float32)
    else:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    kw = dict(resid=out) if epi == K_.EPI_RESID_F32 else {}
    for _ in range(3):
        K_.gemm(a, w, epi, out, cta_group=cg, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        K_.gemm(a, w, epi, out, cta_group=cg, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS for comparison
    o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, w.t(), out=o2)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t(), out=o2)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"   bench gemm M={M} N={N} K={K} cg={cg} epi={epi}: {ms:.3f} ms {tf:.0f} TFLOP/s | cublas {ms2:.3f} ms {2.0*M*N*K/ms2/1e9:.0f} TFLOP/s", flush=True)


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    which = sys.argv[1:] or ["gemm", "attn", "ln", "bench"]
    if "gemm" in which:
        for cg in (1, 2):
            run(f"gemm small cg={cg}", lambda: gemm_case(128 * cg, 256, 64, cg, K.EPI_STORE_F32))
            run(f"gemm k-loop cg={cg}", lambda: gemm_case(128 * cg, 256, 512, cg, K.EPI_STORE_F32))
            run(f"gemm multi-tile cg={cg}", lambda: gemm_case(1024, 1536, 1536, cg, K.EPI_STORE_BF16))
            run(f"gemm ragged cg={cg}", lambda: gemm_case(1000, 384, 48, cg, K.EPI_STORE_BF16))
            run(f"gemm persistent cg={cg}", lambda: gemm_case(12608, 4608, 1536, cg, K.EPI_STORE_BF16))
            run(f"gemm geglu cg={cg}", lambda: gemm_case(1000, 2048, 256, cg, K.EPI_GEGLU_BF16))
            run(f"gemm resid cg={cg}", lambda: gemm_case(1000, 1536, 1024, cg, K.EPI_RESID_F32))
            run(f"gemm gelu cg={cg}", lambda: gemm_case(520, 512, 1536, cg, K.EPI_GELU_BF16))
    if "attn" in which:
        run("attn tiny", lambda: attn_case(2, 17, 4, True, True))
        run("attn 64", lambda: attn_case(2, 64, 4, False, False))
        run("attn 197", lambda: attn_case(3, 197, 24, True, False))
        run("attn 197 pad", lambda: attn_case(3, 197, 4, True, True))
        run("attn 500", lambda: attn_case(2, 500, 4, True, True))
    if "ln" in which:
        run("ln 1536 f32->bf16", lambda: ln_case(1000, 1536, torch.float32, torch.bfloat16, True, False))
        run("ln 6144 bf16->bf16", lambda: ln_case(300, 6144, torch.bfloat16, torch.bfloat16, True, False))
        run("ln 256", lambda: ln_case(77, 256, torch.float32, torch.bfloat16, True, False))
        run("ln 384 gelu merge", lambda: ln_case(2 * 56 * 56, 384, torch.bfloat16, torch.bfloat16, True, True, 56))
        run("ln 512 gelu", lambda: ln_case(999, 512, torch.bfloat16, torch.bfloat16, True, True))
        run("ln 1536 noaffine f32", lambda: ln_case(50, 1536, torch.bfloat16, torch.float32, False, True))
    if "bench" in which:
        for cg in (1, 2):
            run(f"bench qkv cg={cg}", lambda: bench_gemm(12608, 4608, 1536, cg, K.EPI_STORE_BF16))
            run(f"bench geglu cg={cg}", lambda: bench_gemm(12608, 12288, 1536, cg, K.EPI_GEGLU_BF16))
            run(f"bench fc2 cg={cg}", lambda: bench_gemm(12608, 1536, 6144, cg, K.EPI_RESID_F32))


if __name__ == "__main__":
    main()
