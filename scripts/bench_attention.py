"""Times the attention kernel alone at the 4B vision shape (B=64, S=197, H=24) under the current env overrides."""
import os, sys, torch
sys.path.insert(0, ".")
from one_peace_b200 import kernels as K
B, S, H = 64, 197, 24
D = H * 64
qkv = (torch.randn(B * S, 3 * D, device="cuda") * 0.5).bfloat16()
bias = torch.randn(H, S, 200, device="cuda")
out = torch.empty(B * S, D, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    K.attention(qkv, bias, None, B, S, H, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    K.attention(qkv, bias, None, B, S, H, out=out)
e1.record(); torch.cuda.synchronize()
print(f"stage={os.environ.get('OPB_ATTN_STAGE_BIAS')} bpc={os.environ.get('OPB_ATTN_BPC')}: {e0.elapsed_time(e1)/20*1000:.1f} us")
