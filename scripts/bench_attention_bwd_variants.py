"""Where the tcgen05 attention backward spends its time at B=64, S=197, H=24: with / without the bias gather and the dbias
reduction (red.global.add).  usage: python scripts/bench_attention_bwd_variants.py [once]  ('once': one launch per variant, for ncu)"""
import os, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
from one_peace_b200 import kernels as K
B, S, H = 64, 197, 24
D = H * 64
once = len(sys.argv) > 1
g = torch.Generator(device="cuda").manual_seed(S)
qkv = (torch.randn(B * S, 3 * D, device="cuda", generator=g) * 0.6).bfloat16()
s_pad = (S + 3) // 4 * 4
bias = torch.randn(H, S, s_pad, device="cuda", generator=g) * 0.5
d_out = torch.randn(B * S, D, device="cuda", generator=g).bfloat16()
dqkv = torch.zeros(B * S, 3 * D, device="cuda", dtype=torch.bfloat16)
scratch = torch.zeros_like(bias)
for name, bb, db in (("bias+dbias", bias, scratch), ("bias only", bias, None), ("no bias", None, None)):
    lse = torch.empty(B * H * S, device="cuda")
    out = K.attention(qkv, bb, None, B, S, H, lse=lse)
    n = 1 if once else 20
    for _ in range(0 if once else 3):
        K.attention_bwd(qkv, out, d_out, bb, None, lse, dqkv, db, B, S, H, 0.125)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        K.attention_bwd(qkv, out, d_out, bb, None, lse, dqkv, db, B, S, H, 0.125)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / n * 1000:.1f} us (incl. attn_delta)", flush=True)

# transposed bias tables (opb_attention_bwd_t): the form the encoder stack uses
lse = torch.empty(B * H * S, device="cuda")
out = K.attention(qkv, bias, None, B, S, H, lse=lse)
bias_t = K.relpos_bias_transpose(bias)
dbias_t = torch.zeros(H, K.BIAS_T_KEYS, K.BIAS_T_Q, device="cuda")
for name, bt, dt in (("transposed tables: bias_t+dbias_t", bias_t, dbias_t), ("transposed tables: bias_t only", bias_t, None)):
    n = 1 if once else 20
    for _ in range(0 if once else 3):
        K.attention_bwd_t(qkv, out, d_out, bt, None, lse, dqkv, dt, B, S, H, 0.125)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        K.attention_bwd_t(qkv, out, d_out, bt, None, lse, dqkv, dt, B, S, H, 0.125)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / n * 1000:.1f} us (incl. attn_delta)", flush=True)
