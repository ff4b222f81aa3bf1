"""Times the attention kernels alone at the 4B vision shape (B=64, S=197, H=24): mma.sync (dense bias) vs tcgen05 (LUT)."""
import os, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
from one_peace_b200 import kernels as K, relpos
import restated as R
B, S, H = 64, 197, 24
D = H * 64
qkv = (torch.randn(B * S, 3 * D, device="cuda") * 0.5).bfloat16()
bucket = R.make_image_bucket_position(14)
table = torch.randn(732, H, device="cuda")
li = relpos.build_lut_index(bucket.numpy(), relpos.image_codes(S, 14))
rp = K.RelPosBias(lut=K.relpos_lut_build(table, torch.from_numpy(li[0]).cuda()), code_row=torch.from_numpy(li[1]).cuda(), code_col=torch.from_numpy(li[2]).cuda())
dense = K.relpos_bias_build(table, bucket.cuda(), S, H)
out = torch.empty(B * S, D, dtype=torch.bfloat16, device="cuda")
part = torch.empty(H * B * S * 2, device="cuda")
def timeit(f, n=30):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
t_old = timeit(lambda: K.attention(qkv, dense, None, B, S, H, out=out, ln_stats=part))
o_old = out.clone()
os.environ["OPB_ATTN_PERSIST"] = "0"
t_tc = timeit(lambda: K.attention_tc(qkv, rp, None, B, S, H, out=out, ln_stats=part))
o_tc, p_tc = out.clone(), part.clone()
os.environ["OPB_ATTN_PERSIST"] = "1"
t_new = timeit(lambda: K.attention_tc(qkv, rp, None, B, S, H, out=out, ln_stats=part))
print(f"mma.sync {t_old:.1f} us | tcgen05 per-tile CTAs {t_tc:.1f} us | tcgen05 persistent {t_new:.1f} us | max diff vs mma.sync "
      f"{(out.float()-o_old.float()).abs().max().item():.3e} vs per-tile {(out.float()-o_tc.float()).abs().max().item():.3e} "
      f"ln_stats rel diff {((part-p_tc).abs().max()/p_tc.abs().max()).item():.3e}")
# 15 s audio: 750 tokens (adapter/audio.py), 1-D relative positions: mma.sync flash kernel (dense bias) vs two tcgen05 key ranges + merge
Ba, Sa = 16, 750
qa = (torch.randn(Ba * Sa, 3 * D, device="cuda") * 0.5).bfloat16()
ba = R.make_token_bucket_position(256)[:Sa, :Sa]
ta = torch.randn(514, H, device="cuda")
la = relpos.build_lut_index(ba.numpy(), relpos.text_codes(Sa))
rpa = K.RelPosBias(lut=K.relpos_lut_build(ta, torch.from_numpy(la[0]).cuda()), code_row=torch.from_numpy(la[1]).cuda(), code_col=torch.from_numpy(la[2]).cuda())
da = K.relpos_bias_build(ta, ba.cuda(), Sa, H)
oa = torch.empty(Ba * Sa, D, dtype=torch.bfloat16, device="cuda")
t_a0 = timeit(lambda: K.attention(qa, da, None, Ba, Sa, H, out=oa))
o_a0 = oa.clone()
t_a1 = timeit(lambda: K.attention_tc(qa, rpa, None, Ba, Sa, H, out=oa))
print(f"audio B={Ba} S={Sa}: mma.sync {t_a0:.1f} us | tcgen05 two key ranges + merge {t_a1:.1f} us | max diff {(oa.float()-o_a0.float()).abs().max().item():.3e}")
if "timing" in os.environ.get("OPB_LIB_PATH", ""):
    import ctypes
    from one_peace_b200 import _lib
    lib = _lib.load()
    torch.cuda.synchronize()
    lib.opb_attn_timing_dump()
    if hasattr(lib, "opb_tcp_timing_dump"):
        lib.opb_tcp_timing_dump()
