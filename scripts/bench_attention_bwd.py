"""Attention backward at the 4B vision shape (B=64, S=197, H=24) and a text shape: tcgen05 persistent kernel vs the mma.sync pair
(A/B via OPB_ATTN_BWD_TC), results compared with each other and timed with CUDA events."""
import os, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
from one_peace_b200 import kernels as K


def run(B, S, H, pad):
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(S)
    qkv = (torch.randn(B * S, 3 * D, device="cuda", generator=g) * 0.6).bfloat16()
    s_pad = (S + 3) // 4 * 4
    bias = torch.randn(H, S, s_pad, device="cuda", generator=g) * 0.5
    key_pad = None
    if pad:
        key_pad = torch.zeros(B, S, dtype=torch.uint8, device="cuda")
        key_pad[0, S - 5:] = 1
        key_pad[-1, S // 2:] = 1
    d_out = torch.randn(B * S, D, device="cuda", generator=g).bfloat16()
    lse = torch.empty(B * H * S, device="cuda")
    out = K.attention(qkv, bias, key_pad, B, S, H, lse=lse)
    res = {}
    for mode in ("0", "1"):
        os.environ["OPB_ATTN_BWD_TC"] = mode
        dqkv = torch.zeros(B * S, 3 * D, device="cuda", dtype=torch.bfloat16)
        dbias = torch.zeros(H, S, s_pad, device="cuda")
        K.attention_bwd(qkv, out, d_out, bias, key_pad, lse, dqkv, dbias, B, S, H, 0.125)
        torch.cuda.synchronize()
        scratch = torch.zeros_like(dbias)
        for _ in range(3):
            K.attention_bwd(qkv, out, d_out, bias, key_pad, lse, dqkv, scratch, B, S, H, 0.125)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            K.attention_bwd(qkv, out, d_out, bias, key_pad, lse, dqkv, scratch, B, S, H, 0.125)
        e1.record(); torch.cuda.synchronize()
        res[mode] = (dqkv.float(), dbias.clone(), e0.elapsed_time(e1) / 20 * 1000)
    a, b = res["0"], res["1"]
    rel = lambda x, y: ((x - y).abs().max() / (y.abs().max() + 1e-9)).item()
    print(f"B={B} S={S} H={H} pad={pad}: mma.sync {a[2]:.1f} us | tcgen05 {b[2]:.1f} us | rel diff dq {rel(b[0][:, :D], a[0][:, :D]):.3e} "
          f"dk {rel(b[0][:, D:2*D], a[0][:, D:2*D]):.3e} dv {rel(b[0][:, 2*D:], a[0][:, 2*D:]):.3e} dbias {rel(b[1], a[1]):.3e}", flush=True)


for cfg in ((2, 33, 2, True), (3, 197, 4, False), (2, 100, 2, True), (2, 214, 3, True), (64, 197, 24, False), (64, 33, 24, True)):
    run(*cfg)
