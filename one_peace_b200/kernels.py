"""Tensor-level wrappers over the C-ABI: PyTorch allocates, the sm_100a kernels compute.

Every function takes CUDA tensors, passes raw device pointers + the current stream to
``libonepeace_b200.so`` and returns torch tensors.  Nothing here computes with torch ops.
"""
import torch

from . import _lib

F32, BF16 = 0, 1
EPI_STORE_BF16, EPI_GEGLU_BF16, EPI_RESID_F32, EPI_STORE_F32, EPI_GELU_BF16 = 0, 1, 2, 3, 4


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("one_peace_b200 kernels need CUDA tensors (there is no CPU path)")


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError(f"unsupported dtype {t.dtype}")


def gemm(a, w, epi, out, *, bias=None, colscale=None, gamma=None, resid=None, out_group=0, out_group_stride=0,
         out_row_offset=0, resid_period=0, resid_row_offset=0, cta_group=0, M=None, lda=None, K=None):
    """out = epilogue(a[M,K] @ w[N,K]^T).  a/w bf16; `lda`/`M`/`K` allow strided (even overlapping) row views."""
    _need_cuda(a, w, out)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    if M is None:
        M = a.shape[0]
    if K is None:
        K = a.shape[1]
    if lda is None:
        lda = a.stride(0)
    N = w.shape[0]
    assert w.shape[1] == K and w.stride(1) == 1 and a.stride(-1) == 1
    ldr = resid.stride(-2) if resid is not None else 0
    st = _lib.load().opb_gemm_bf16(a.data_ptr(), lda, w.data_ptr(), w.stride(0), M, N, K, epi, out.data_ptr(),
                                   out.stride(-2), _ptr(bias), _ptr(colscale), _ptr(gamma), _ptr(resid), ldr,
                                   out_group, out_group_stride, out_row_offset, resid_period, resid_row_offset,
                                   cta_group, _stream())
    _lib.check(st, "opb_gemm_bf16")
    return out


def attention(qkv, bias, key_pad, B, S, H, out=None, lse=None):
    _need_cuda(qkv, bias, key_pad)
    D = H * 64
    assert qkv.dtype == torch.bfloat16 and qkv.shape == (B * S, 3 * D) and qkv.is_contiguous()
    if out is None:
        out = torch.empty(B * S, D, dtype=torch.bfloat16, device=qkv.device)
    s_pad = 0
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.shape[0] == H and bias.shape[1] == S
        s_pad = bias.shape[2]
    if key_pad is not None:
        assert key_pad.dtype == torch.uint8 and key_pad.shape == (B, S) and key_pad.is_contiguous()
    st = _lib.load().opb_attention_fwd(qkv.data_ptr(), _ptr(bias), _ptr(key_pad), out.data_ptr(), _ptr(lse), B, S,
                                       H, s_pad, _stream())
    _lib.check(st, "opb_attention_fwd")
    return out


def layernorm(x, gamma, beta, out, *, rows=None, dim=None, ld_in=None, ld_out=None, eps=1e-5, gelu=False,
              merge_grid_w=0):
    _need_cuda(x, out)
    if rows is None:
        rows = x.shape[0]
    if dim is None:
        dim = x.shape[-1]
    if ld_in is None:
        ld_in = x.stride(-2)
    if ld_out is None:
        ld_out = out.stride(-2)
    st = _lib.load().opb_layernorm(x.data_ptr(), _dt(x), ld_in, out.data_ptr(), _dt(out), ld_out, _ptr(gamma),
                                   _ptr(beta), rows, dim, eps, int(gelu), merge_grid_w, _stream())
    _lib.check(st, "opb_layernorm")
    return out
