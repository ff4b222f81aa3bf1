"""Tensor-level wrappers over the C-ABI: PyTorch allocates, the sm_100a kernels compute.

Every function takes CUDA tensors, passes raw device pointers + the current stream to
``libonepeace_b200.so`` and returns torch tensors.  Nothing here computes with torch ops.
"""
import torch

from . import _lib

F32, BF16 = 0, 1
EPI_STORE_BF16, EPI_GEGLU_BF16, EPI_RESID_F32, EPI_STORE_F32, EPI_GELU_BF16 = 0, 1, 2, 3, 4


# number of kernel launches issued through the C-ABI since import (bench.py reports it per step)
LAUNCHES = 0
# optional per-call profiler hook: bench.py installs a callable(name, flops) -> context manager
PROFILE_HOOK = None


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


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
         out_row_offset=0, out_group_valid=0, resid_period=0, resid_row_offset=0, cta_group=0, M=None, lda=None,
         K=None):
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
    if PROFILE_HOOK is not None:
        PROFILE_HOOK("gemm_begin", 0.0, None)
    assert w.shape[1] == K and w.stride(1) == 1 and a.stride(-1) == 1
    ldr = resid.stride(-2) if resid is not None else 0
    st = _lib.load().opb_gemm_bf16(a.data_ptr(), lda, w.data_ptr(), w.stride(0), M, N, K, epi, out.data_ptr(),
                                   out.stride(-2), _ptr(bias), _ptr(colscale), _ptr(gamma), _ptr(resid), ldr,
                                   out_group, out_group_stride, out_row_offset, out_group_valid, resid_period,
                                   resid_row_offset, cta_group, _stream())
    _lib.check(st, "opb_gemm_bf16")
    _count()
    if PROFILE_HOOK is not None:
        PROFILE_HOOK("gemm", 2.0 * M * N * K, (M, N, K, epi))
    return out


class RelPosBias:
    """Relative-position bias of one forward in both forms the attention kernels take: the dense fp32 (H,S,S_pad)
    table (mma.sync kernel, any S) and the LUT form (tcgen05 kernels, S <= 768)."""

    def __init__(self, dense=None, lut=None, code_row=None, code_col=None, seg_split=0):
        self.dense, self.lut, self.code_row, self.code_col, self.seg_split = dense, lut, code_row, code_col, seg_split
        self.lut_max = None


def build_segmented_lut(parts, device):
    """Concatenated ('vl' / 'al') LUT-form bias for `attention_tc`.  parts = [(table fp32 [NB,H], lut_index, n)] per modality
    in sequence order, lut_index = relpos.build_lut_index(...) result (numpy lut_idx, code_row, code_col) or device tensors.
    The per-modality LUTs are laid end to end; row codes of the second modality are shifted by the length of the first LUT
    so that same-modality code differences land in that modality's LUT; the kernel zeroes the cross-modality bias."""
    assert len(parts) == 2, "two concatenated modalities (transformer_encoder.py:127-134)"
    luts, rows, cols, off = [], [], [], 0
    for table, li, n in parts:
        idx, cr, cc = (torch.as_tensor(a, device=device).to(torch.int32) for a in li)
        luts.append(relpos_lut_build(table, idx.contiguous()))
        rows.append(cr[:n] + off)
        cols.append(cc[:n])
        off += idx.numel()
    lut = torch.cat(luts, dim=1).contiguous()
    pad4 = lambda t: torch.cat([t, torch.zeros((-t.numel()) % 4, dtype=t.dtype, device=t.device)]).contiguous()
    return RelPosBias(lut=lut, code_row=pad4(torch.cat(rows)), code_col=pad4(torch.cat(cols)), seg_split=parts[0][2])


def relpos_lut_build(table, idx):
    """table fp32 [NB,H], idx int32 [L] -> lut fp32 [H, L]"""
    L, H = idx.numel(), table.shape[1]
    lut = torch.empty(H, L, dtype=torch.float32, device=table.device)
    st = _lib.load().opb_relpos_lut_build(table.data_ptr(), idx.data_ptr(), lut.data_ptr(), L, H, _stream())
    _lib.check(st, "opb_relpos_lut_build")
    _count()
    return lut


# tcgen05 attention forward: persistent kernel S <= 224, per-tile kernel S <= 384.  384 < S <= 768 also runs on tcgen05 (two key
# ranges + merge, csrc/attention_tc.cu) but measured SLOWER than the flash-style mma.sync kernel at the audio shape (B = 16,
# S = 750, H = 24: 601 us vs 465 us, profiles/r02_attention_fwd_bwd.txt): with 384 keys per launch the per-tile kernel runs one CTA
# per SM and its TMA / MMA / soft-max phases no longer overlap.  The adapters therefore keep the dense-bias kernel for S > 384 unless
# OPB_ATTN_LONG_TC=1.
ATTN_TC_MAX_S = 768 if __import__("os").environ.get("OPB_ATTN_LONG_TC", "0") == "1" else 384


def attention_tc(qkv, rp, key_pad, B, S, H, out=None, ln_stats=None, lse=None):
    """tcgen05 attention (S <= 768).  rp: RelPosBias with the LUT form (rp.seg_split > 0: two concatenated modalities,
    block-diagonal bias; S <= 384 then)."""
    D = H * 64
    assert qkv.dtype == torch.bfloat16 and qkv.shape == (B * S, 3 * D) and qkv.is_contiguous()
    if out is None:
        out = torch.empty(B * S, D, dtype=torch.bfloat16, device=qkv.device)
    if getattr(rp, "lut_max", None) is None:        # per-head bound of the bias (once per table build)
        mx = rp.lut.amax(dim=1)
        rp.lut_max = (mx.clamp_min(0.0) if int(getattr(rp, "seg_split", 0)) > 0 else mx).contiguous()
    st = _lib.load().opb_attention_tc_fwd(qkv.data_ptr(), rp.lut.data_ptr(), rp.lut_max.data_ptr(), rp.lut.shape[1], rp.code_row.data_ptr(),
                                          rp.code_col.data_ptr(), _ptr(key_pad), out.data_ptr(), _ptr(lse), _ptr(ln_stats), B, S,
                                          H, int(getattr(rp, "seg_split", 0)), _stream())
    _lib.check(st, "opb_attention_tc_fwd")
    _count()
    return out


def gemm_ln(a, w, epi, out, *, ln_mu=None, ln_rstd=None, ln_colsum=None, bias=None, colscale=None, gamma=None,
            resid=None, stats_out=None, out_bf16=None, cta_group=0, workspace=None, ln_partial=None):
    """GEMM through `opb_gemm_bf16_ex`: fused LayerNorm of the A operand (ln_*), statistics / bf16 side outputs."""
    _need_cuda(a, w, out)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.stride(-1) == 1 and w.stride(1) == 1
    M, Kd = a.shape
    N = w.shape[0]
    args = _lib.GemmArgs()
    args.A, args.lda, args.B, args.ldb = a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0)
    args.M, args.N, args.K, args.epi = M, N, Kd, epi
    args.out, args.ldo = out.data_ptr(), out.stride(-2)
    args.bias, args.colscale, args.gamma, args.resid = _ptr(bias) or None, _ptr(colscale) or None, _ptr(gamma) or None, _ptr(resid) or None
    args.ldr = resid.stride(-2) if resid is not None else 0
    args.ln_mu, args.ln_rstd, args.ln_colsum = _ptr(ln_mu) or None, _ptr(ln_rstd) or None, _ptr(ln_colsum) or None
    args.stats_out = _ptr(stats_out) or None
    args.out_bf16 = _ptr(out_bf16) or None
    args.ldo_bf16 = out_bf16.stride(-2) if out_bf16 is not None else 0
    args.cta_group = cta_group
    if ln_partial is not None:      # (records tensor, parts, dim, eps)
        args.ln_partial, args.ln_parts, args.ln_dim, args.ln_eps = ln_partial[0].data_ptr(), ln_partial[1], ln_partial[2], ln_partial[3]
    if workspace is not None:
        args.workspace, args.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    if PROFILE_HOOK is not None:
        PROFILE_HOOK("gemm_begin", 0.0, None)
    import ctypes as _ct
    st = _lib.load().opb_gemm_bf16_ex(_ct.addressof(args), _stream())
    _lib.check(st, "opb_gemm_bf16_ex")
    _count()
    if PROFILE_HOOK is not None:
        PROFILE_HOOK("gemm", 2.0 * M * N * Kd, (M, N, Kd, epi))
    return out


def row_stats_cast(x, out_bf16, mu, rstd, eps=1e-5):
    rows, dim = x.shape
    st = _lib.load().opb_row_stats_cast(x.data_ptr(), x.stride(0), out_bf16.data_ptr(), out_bf16.stride(0), mu.data_ptr(),
                                        rstd.data_ptr(), rows, dim, eps, _stream())
    _lib.check(st, "opb_row_stats_cast")
    _count()


def ln_stats_finalize(partial, parts, rows, dim, eps, mu, rstd):
    st = _lib.load().opb_ln_stats_finalize(partial.data_ptr(), parts, rows, dim, eps, mu.data_ptr(), rstd.data_ptr(),
                                           _stream())
    _lib.check(st, "opb_ln_stats_finalize")
    _count()


def attention(qkv, bias, key_pad, B, S, H, out=None, lse=None, ln_stats=None):
    """mma.sync attention, any S.  bias: dense fp32 (H,S,S_pad) shared by the batch, or (B,H,S,S_pad) per sample."""
    _need_cuda(qkv, bias, key_pad)
    D = H * 64
    assert qkv.dtype == torch.bfloat16 and qkv.shape == (B * S, 3 * D) and qkv.is_contiguous()
    if out is None:
        out = torch.empty(B * S, D, dtype=torch.bfloat16, device=qkv.device)
    s_pad, bstride = 0, 0
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.shape[-3] == H and bias.shape[-2] == S
        s_pad = bias.shape[-1]
        if bias.dim() == 4:
            assert bias.shape[0] == B
            bstride = H * S * s_pad
    if key_pad is not None:
        assert key_pad.dtype == torch.uint8 and key_pad.shape == (B, S) and key_pad.is_contiguous()
    st = _lib.load().opb_attention_fwd(qkv.data_ptr(), _ptr(bias), _ptr(key_pad), out.data_ptr(), _ptr(lse),
                                       _ptr(ln_stats), B, S, H, s_pad, bstride, _stream())
    _lib.check(st, "opb_attention_fwd")
    _count()
    return out


def layernorm(x, gamma, beta, out, *, rows=None, dim=None, ld_in=None, ld_out=None, eps=1e-5, gelu=False,
              merge_grid_w=0, row_period=0, row_valid=0, out_period=0, out_row_shift=0, group_in=0, group_out=0,
              accumulate=False):
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
                                   _ptr(beta), rows, dim, eps, int(gelu), merge_grid_w, row_period, row_valid,
                                   out_period, out_row_shift, group_in, group_out, int(accumulate), _stream())
    _lib.check(st, "opb_layernorm")
    _count()
    return out


def grouped_conv1d(x_halo, w, bias, out, rows, groups, c_pad, taps, n_per_group, epi=EPI_STORE_BF16):
    """out[r, g*n+co] = bias + sum_j sum_c x_halo[r+j, g, c] * w[g*n+co, j*c_pad+c]  (see onepeace_b200.h)."""
    _need_cuda(x_halo, w, out)
    assert x_halo.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x_halo.is_contiguous() and w.is_contiguous()
    assert x_halo.numel() >= (rows + taps - 1) * groups * c_pad and w.shape == (groups * n_per_group, taps * c_pad)
    st = _lib.load().opb_grouped_conv1d_bf16(x_halo.data_ptr(), w.data_ptr(), rows, groups, c_pad, taps, n_per_group,
                                             epi, out.data_ptr(), out.stride(-2), _ptr(bias), _stream())
    _lib.check(st, "opb_grouped_conv1d_bf16")
    _count()
    return out


def pack_group_halo(x, out, B, T, x_period, x_row_shift, out_period, halo, dim, group_in, group_out):
    _need_cuda(x, out)
    assert x.dtype == torch.float32 and out.dtype == torch.bfloat16
    st = _lib.load().opb_pack_group_halo(x.data_ptr(), x.stride(-2), out.data_ptr(), B, T, x_period, x_row_shift,
                                         out_period, halo, dim, group_in, group_out, _stream())
    _lib.check(st, "opb_pack_group_halo")
    _count()
    return out


def text_embed(tokens, table, pos, cls, pad_idx=1):
    """-> (x fp32 [B,T+1,D], pad_mask uint8 [B,T+1])"""
    _need_cuda(tokens, table, pos, cls)
    B, T = tokens.shape
    D = table.shape[1]
    assert tokens.dtype == torch.int64 and tokens.is_contiguous() and table.is_contiguous()
    assert pos.dtype == torch.float32 and cls.dtype == torch.float32 and pos.shape[0] >= T + 1
    x = torch.empty(B, T + 1, D, dtype=torch.float32, device=tokens.device)
    pad = torch.empty(B, T + 1, dtype=torch.uint8, device=tokens.device)
    st = _lib.load().opb_text_embed(tokens.data_ptr(), table.data_ptr(), _dt(table), pos.data_ptr(), cls.data_ptr(),
                                    x.data_ptr(), pad.data_ptr(), B, T, D, pad_idx, _stream())
    _lib.check(st, "opb_text_embed")
    _count()
    return x, pad


def image_patchify4(img):
    _need_cuda(img)
    B, C, R, R2 = img.shape
    assert C == 3 and R == R2 and img.is_contiguous()
    out = torch.empty(B * (R // 4) * (R // 4), 48, dtype=torch.bfloat16, device=img.device)
    st = _lib.load().opb_image_patchify4(img.data_ptr(), _dt(img), out.data_ptr(), B, R, _stream())
    _lib.check(st, "opb_image_patchify4")
    _count()
    return out


def cls_row_init(cls, pos0, x):
    """x fp32 [B,S,D]: x[:,0,:] = cls + pos0"""
    _need_cuda(cls, pos0, x)
    B, S, D = x.shape
    st = _lib.load().opb_cls_row_init(cls.data_ptr(), pos0.data_ptr(), x.data_ptr(), S * D, B, D, _stream())
    _lib.check(st, "opb_cls_row_init")
    _count()
    return x


def relpos_bias_build(table, bucket, S, H):
    """table fp32 [NB,H], bucket int64 [R,R] -> fp32 [H,S,s_pad]"""
    _need_cuda(table, bucket)
    assert table.dtype == torch.float32 and table.is_contiguous() and table.shape[1] == H
    assert bucket.dtype == torch.int64 and bucket.stride(1) == 1 and bucket.shape[0] >= S and bucket.shape[1] >= S
    s_pad = (S + 7) // 8 * 8
    bias = torch.empty(H, S, s_pad, dtype=torch.float32, device=table.device)
    st = _lib.load().opb_relpos_bias_build(table.data_ptr(), bucket.data_ptr(), bias.data_ptr(), S, s_pad, H,
                                           bucket.stride(0), _stream())
    _lib.check(st, "opb_relpos_bias_build")
    _count()
    return bias


def audio_frame10(wav, pitch, out):
    _need_cuda(wav, out)
    B, N = wav.shape
    assert wav.is_contiguous()
    st = _lib.load().opb_audio_frame10(wav.data_ptr(), _dt(wav), out.data_ptr(), B, N, pitch, _stream())
    _lib.check(st, "opb_audio_frame10")
    _count()
    return out


def l2_normalize_rows(x, want_bf16=False):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.stride(1) == 1
    rows, D = x.shape
    y = torch.empty(rows, D, dtype=torch.float32, device=x.device)
    y16 = torch.empty(rows, D, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    st = _lib.load().opb_l2_normalize_rows(x.data_ptr(), x.stride(0), y.data_ptr(), _ptr(y16), rows, D, _stream())
    _lib.check(st, "opb_l2_normalize_rows")
    _count()
    return (y, y16) if want_bf16 else y


def zero_padded_rows(x, pad_mask):
    _need_cuda(x, pad_mask)
    D = x.shape[-1]
    rows = x.numel() // D
    assert x.dtype == torch.float32 and x.is_contiguous() and pad_mask.dtype == torch.uint8 and pad_mask.numel() == rows
    st = _lib.load().opb_zero_padded_rows(x.data_ptr(), pad_mask.data_ptr(), rows, D, _stream())
    _lib.check(st, "opb_zero_padded_rows")
    _count()
    return x


# ----------------------------------------------------------------------------------------------------
# contrastive head
# ----------------------------------------------------------------------------------------------------
def transpose_bf16(x, rows=None, cols=None):
    """bf16 [rows, cols] view (row pitch x.stride(0)) -> contiguous [cols, rows]"""
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.stride(1) == 1 and x.dim() == 2
    rows = x.shape[0] if rows is None else rows
    cols = x.shape[1] if cols is None else cols
    out = torch.empty(cols, rows, dtype=torch.bfloat16, device=x.device)
    st = _lib.load().opb_transpose_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), rows, cols, _stream())
    _lib.check(st, "opb_transpose_bf16")
    _count()
    return out


def split_bf16x3(x, side):
    """fp32 [r,d] -> bf16 [r,3d]; side 0 = [hi|hi|lo] (local operand), side 1 = [hi|lo|hi] (gathered operand)."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    r, d = x.shape
    out = torch.empty(r, 3 * d, dtype=torch.bfloat16, device=x.device)
    st = _lib.load().opb_split_bf16x3(x.data_ptr(), out.data_ptr(), r, d, side, _stream())
    _lib.check(st, "opb_split_bf16x3")
    _count()
    return out


def split_bf16x3_x4(xs, sides):
    """four fp32 [r_i, d] tensors -> four bf16 [r_i, 3d] splits in ONE launch (the operands of one InfoNCE step)"""
    import ctypes
    assert len(xs) == 4 and len(sides) == 4
    d = xs[0].shape[1]
    for x in xs:
        _need_cuda(x)
        assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2 and x.shape[1] == d
    outs = [torch.empty(x.shape[0], 3 * d, dtype=torch.bfloat16, device=x.device) for x in xs]
    px = (ctypes.c_void_p * 4)(*[x.data_ptr() for x in xs])
    po = (ctypes.c_void_p * 4)(*[o.data_ptr() for o in outs])
    pr = (ctypes.c_int64 * 4)(*[x.shape[0] for x in xs])
    ps = (ctypes.c_int * 4)(*[int(v) for v in sides])
    st = _lib.load().opb_split_bf16x3_x4(ctypes.cast(px, ctypes.c_void_p), ctypes.cast(po, ctypes.c_void_p),
                                         ctypes.cast(pr, ctypes.c_void_p), ctypes.cast(ps, ctypes.c_void_p), d, _stream())
    _lib.check(st, "opb_split_bf16x3_x4")
    _count()
    return outs


_TICKETS = {}


def infonce_forward2(a3, b3, a_all3, b_all3, scale, target_offset, eps, n_valid=0):
    """Both directions of the InfoNCE forward in 3 launches (two LSE_PARTIAL GEMMs + one merge / reduce kernel).
    -> (lse_a [b], lse_b [b], out3 = {loss, #correct a->b, #correct b->a})"""
    lib = _lib.load()
    b, k = a3.shape
    n = a_all3.shape[0]
    dev = a3.device
    ws_a = torch.empty(lib.opb_infonce_ws_floats(b, n), dtype=torch.float32, device=dev)
    ws_b = torch.empty_like(ws_a)
    for x, y, ws in ((a3, b_all3, ws_a), (b3, a_all3, ws_b)):
        st = lib.opb_infonce_lse_gemm(x.data_ptr(), y.data_ptr(), scale.data_ptr(), b, n, k, target_offset, ws.data_ptr(), int(n_valid),
                                      _stream())
        _lib.check(st, "opb_infonce_lse_gemm")
    lse_a = torch.empty(b, dtype=torch.float32, device=dev)
    lse_b = torch.empty(b, dtype=torch.float32, device=dev)
    loss_ab = torch.empty(2 * b, dtype=torch.float32, device=dev)
    am_ab = torch.empty(2 * b, dtype=torch.int32, device=dev)
    out3 = torch.empty(3, dtype=torch.float32, device=dev)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _TICKETS:
        _TICKETS[key] = torch.zeros(1, dtype=torch.int32, device=dev)       # the kernel leaves it at zero
    st = lib.opb_infonce_merge_reduce(ws_a.data_ptr(), ws_b.data_ptr(), b, n, int(n_valid), float(eps), target_offset, lse_a.data_ptr(),
                                      lse_b.data_ptr(), loss_ab.data_ptr(), am_ab.data_ptr(), out3.data_ptr(), _TICKETS[key].data_ptr(),
                                      _stream())
    _lib.check(st, "opb_infonce_merge_reduce")
    _count(3)
    return lse_a, lse_b, out3


def infonce_rows(a_local, b_all, scale, target_offset, eps, n_valid=0):
    """One direction of the InfoNCE forward.  a_local bf16 [b,k], b_all bf16 [n,k], scale fp32 device scalar.
    n_valid > 0: only the first n_valid rows of b_all are classes (the rest is zero padding to n % 8 == 0).
    -> (row_lse [b], row_loss [b], row_argmax int32 [b])"""
    _need_cuda(a_local, b_all, scale)
    b, d = a_local.shape
    n = b_all.shape[0]
    assert a_local.dtype == torch.bfloat16 and b_all.dtype == torch.bfloat16 and scale.dtype == torch.float32
    assert a_local.is_contiguous() and b_all.is_contiguous() and b_all.shape[1] == d
    lib = _lib.load()
    dev = a_local.device
    ws = torch.empty(lib.opb_infonce_ws_floats(b, n), dtype=torch.float32, device=dev)
    lse = torch.empty(b, dtype=torch.float32, device=dev)
    loss = torch.empty(b, dtype=torch.float32, device=dev)
    amax = torch.empty(b, dtype=torch.int32, device=dev)
    st = lib.opb_infonce_rows(a_local.data_ptr(), b_all.data_ptr(), scale.data_ptr(), b, n, d, target_offset, eps,
                              ws.data_ptr(), lse.data_ptr(), loss.data_ptr(), amax.data_ptr(), int(n_valid), _stream())
    _lib.check(st, "opb_infonce_rows")
    _count(2)
    return lse, loss, amax


def infonce_reduce(loss_a, loss_b, am_a, am_b, target_offset):
    out = torch.empty(3, dtype=torch.float32, device=loss_a.device)
    st = _lib.load().opb_infonce_reduce(loss_a.data_ptr(), loss_b.data_ptr(), am_a.data_ptr(), am_b.data_ptr(),
                                        loss_a.numel(), target_offset, out.data_ptr(), _stream())
    _lib.check(st, "opb_infonce_reduce")
    _count()
    return out


def infonce_grad(a_local, b_all, bT_all, scale, row_lse, target_offset, eps, n_valid=0, coef=0.0, d=None):
    """-> (grad_a fp32 [b,d], ws_gz) for one direction; a_local/b_all [.,k] (k = d or 3d).  bT_all: bf16 [d,n] transposed copy
    of b_all's first d columns, or None (then pass d): b_all is read in place as an MN-major operand.
    coef = weight of one row's loss (0 -> 1 / (2 b), the two-direction InfoNCE mean)"""
    b, k = a_local.shape
    n = b_all.shape[0]
    d = bT_all.shape[0] if bT_all is not None else int(d)
    dev = a_local.device
    g_ws = torch.empty(b, n, dtype=torch.bfloat16, device=dev)
    ws_gz = torch.empty((n + 255) // 256 * b, dtype=torch.float32, device=dev)
    grad = torch.empty(b, d, dtype=torch.float32, device=dev)
    st = _lib.load().opb_infonce_grad(a_local.data_ptr(), b_all.data_ptr(), _ptr(bT_all), scale.data_ptr(),
                                      row_lse.data_ptr(), b, n, d, k, target_offset, eps, g_ws.data_ptr(),
                                      ws_gz.data_ptr(), grad.data_ptr(), int(n_valid), float(coef), _stream())
    _lib.check(st, "opb_infonce_grad")
    _count(2)
    return grad, ws_gz


def infonce_dscale(ws_a, ws_b, b, n):
    out = torch.empty(1, dtype=torch.float32, device=ws_a.device)
    st = _lib.load().opb_infonce_dscale(ws_a.data_ptr(), ws_b.data_ptr(), b, n, out.data_ptr(), _stream())
    _lib.check(st, "opb_infonce_dscale")
    _count()
    return out


# ----------------------------------------------------------------------------------------------------------------
# backward pass
# ----------------------------------------------------------------------------------------------------------------
_BWD_WS = {}


def bwd_ws(dim, device):
    """fp32 scratch for the column-reduction kernels (per device, grown on demand)."""
    need = _lib.load().opb_bwd_ws_floats(int(dim))
    key = (device.index if device.index is not None else torch.cuda.current_device())
    ws = _BWD_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.float32, device=device)
        _BWD_WS[key] = ws
    return ws


def layernorm_bwd(x, dy, gamma, beta, dx, *, eps=1e-5, gelu=False, accumulate=False, dgamma=None, dbeta=None, rows=None,
                  dim=None, ldx=None, ld_dx=None, dy_merge_w=0):
    """x, dy, dx: [rows, dim] (fp32 / bf16, row stride free); dgamma / dbeta fp32 [dim] outputs (optional)."""
    _need_cuda(x, dy, dx)
    rows = x.shape[0] if rows is None else rows
    dim = x.shape[-1] if dim is None else dim
    ws = bwd_ws(dim, x.device) if (dgamma is not None or dbeta is not None) else None
    st = _lib.load().opb_layernorm_bwd(x.data_ptr(), _dt(x), x.stride(-2) if ldx is None else ldx, dy.data_ptr(), _dt(dy),
                                       dy.stride(-2), _ptr(gamma), _ptr(beta), dx.data_ptr(), _dt(dx),
                                       dx.stride(-2) if ld_dx is None else ld_dx, int(accumulate), rows, dim, eps,
                                       int(gelu), dy_merge_w, _ptr(ws), _ptr(dgamma), _ptr(dbeta), _stream())
    _lib.check(st, "opb_layernorm_bwd")
    _count(1 + (dgamma is not None) + (dbeta is not None))
    return dx


def geglu_fwd(gl, u):
    rows, F2 = gl.shape
    st = _lib.load().opb_geglu_fwd(gl.data_ptr(), u.data_ptr(), rows, F2 // 2, _stream())
    _lib.check(st, "opb_geglu_fwd")
    _count()
    return u


def geglu_bwd(gl, du, dgl):
    rows, F2 = gl.shape
    st = _lib.load().opb_geglu_bwd(gl.data_ptr(), du.data_ptr(), dgl.data_ptr(), rows, F2 // 2, _stream())
    _lib.check(st, "opb_geglu_bwd")
    _count()
    return dgl


def scale_resid_fwd(x, o, gamma, row_scale, out):
    rows, n = x.shape
    st = _lib.load().opb_scale_resid_fwd(x.data_ptr(), o.data_ptr(), _ptr(gamma), _ptr(row_scale), out.data_ptr(), rows, n,
                                         _stream())
    _lib.check(st, "opb_scale_resid_fwd")
    _count()
    return out


def scale_resid_bwd(dx, o, gamma, row_scale, d_o, dgamma=None, dbias=None, in_period=0, in_valid=0, in_shift=0):
    """d_o bf16 [rows, n] = row_scale * gamma * dx (rows of dx optionally gathered, see onepeace_b200.h)."""
    rows, n = d_o.shape
    ws = bwd_ws(n, dx.device)
    st = _lib.load().opb_scale_resid_bwd(dx.data_ptr(), _ptr(o), _ptr(gamma), _ptr(row_scale), d_o.data_ptr(), ws.data_ptr(),
                                         _ptr(dgamma), _ptr(dbias), rows, n, in_period, in_valid, in_shift, _stream())
    _lib.check(st, "opb_scale_resid_bwd")
    _count(1 + (dgamma is not None) + (dbias is not None))
    return d_o


def colsum(y, out):
    rows, n = y.shape
    ws = bwd_ws(n, y.device)
    st = _lib.load().opb_colsum_bf16(y.data_ptr(), y.stride(0), ws.data_ptr(), out.data_ptr(), rows, n, _stream())
    _lib.check(st, "opb_colsum_bf16")
    _count(2)
    return out


def attention_bwd(qkv, out, d_out, bias, key_pad, lse, dqkv, dbias, B, S, H, q_scale):
    """bias / dbias: dense fp32 (H,S,S_pad) tables shared by the batch, (B,H,S,S_pad) per-sample tables, or None;
    lse fp32 [B,H,S] from `attention(..., lse=)`."""
    delta = torch.empty(B * H * S, dtype=torch.float32, device=qkv.device)
    s_pad = bias.shape[-1] if bias is not None else 0
    bstride = H * S * s_pad if (bias is not None and bias.dim() == 4) else 0
    if dbias is not None:
        assert dbias.shape == bias.shape
    st = _lib.load().opb_attention_bwd(qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), _ptr(bias), _ptr(key_pad),
                                       lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), _ptr(dbias), B, S, H, s_pad,
                                       float(q_scale), bstride, _stream())
    _lib.check(st, "opb_attention_bwd")
    _count(3)
    return dqkv


BIAS_T_KEYS, BIAS_T_Q = 256, 224      # transposed bias tables of the tcgen05 attention backward (csrc/attention_bwd_tc.cu)


def relpos_bias_transpose(bias):
    """dense fp32 (H,S,S_pad) bias -> (H,256,112) half2 words (int32 storage), x log2 e, zero-padded; S <= 224."""
    H, S, s_pad = bias.shape
    out = torch.empty(H, BIAS_T_KEYS, BIAS_T_Q // 2, dtype=torch.int32, device=bias.device)
    st = _lib.load().opb_relpos_bias_transpose(bias.data_ptr(), out.data_ptr(), S, s_pad, H, _stream())
    _lib.check(st, "opb_relpos_bias_transpose")
    _count()
    return out


def relpos_dbias_fold(dbias_t, dbias):
    """dbias fp32 (H,S,S_pad) += transpose of dbias_t fp32 (H,256,224)"""
    H, S, s_pad = dbias.shape
    st = _lib.load().opb_relpos_dbias_fold(dbias_t.data_ptr(), dbias.data_ptr(), S, s_pad, H, _stream())
    _lib.check(st, "opb_relpos_dbias_fold")
    _count()
    return dbias


def relpos_dbias_center(dbias):
    """in place: every row of the dense fp32 (H,S,S_pad) bias gradient gets its mean over the S valid columns subtracted"""
    H, S, s_pad = dbias.shape
    st = _lib.load().opb_relpos_dbias_center(dbias.data_ptr(), S, s_pad, H, _stream())
    _lib.check(st, "opb_relpos_dbias_center")
    _count()
    return dbias


def attention_bwd_t(qkv, out, d_out, bias_t, key_pad, lse, dqkv, dbias_t, B, S, H, q_scale):
    """tcgen05 attention backward with transposed bias tables (relpos_bias_transpose / a zeroed (H,256,224) fp32 dbias_t that
    several layers may share); S <= 224."""
    delta = torch.empty(B * H * S, dtype=torch.float32, device=qkv.device)
    st = _lib.load().opb_attention_bwd_t(qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), _ptr(bias_t), _ptr(key_pad),
                                         lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), _ptr(dbias_t), B, S, H,
                                         float(q_scale), _stream())
    _lib.check(st, "opb_attention_bwd_t")
    _count(2)
    return dqkv


def relpos_bias_bwd(dbias, bucket, dtable, S):
    H, s_pad = dbias.shape[0], dbias.shape[-1]
    st = _lib.load().opb_relpos_bias_bwd(dbias.data_ptr(), bucket.data_ptr(), dtable.data_ptr(), S, s_pad, H, bucket.stride(0),
                                         _stream())
    _lib.check(st, "opb_relpos_bias_bwd")
    _count()
    return dtable


def batch_sum(x, out, B, n, ld, accumulate=False):
    """out[c] (+)= sum_b x[b * ld + c], c < n (x is addressed through its data pointer: pass a view of the first row)"""
    st = _lib.load().opb_batch_sum_f32(x.data_ptr(), ld, out.data_ptr(), B, n, int(accumulate), _stream())
    _lib.check(st, "opb_batch_sum_f32")
    _count()
    return out


def l2_normalize_bwd(x, dy, want_f32=False):
    """x fp32 [rows, D] (the un-normalised rows), dy fp32 -> bf16 dx (and fp32 when asked)"""
    rows, D = x.shape
    dx16 = torch.empty(rows, D, dtype=torch.bfloat16, device=x.device)
    dx32 = torch.empty(rows, D, dtype=torch.float32, device=x.device) if want_f32 else None
    st = _lib.load().opb_l2_normalize_bwd(x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), _ptr(dx32), dx16.data_ptr(),
                                          rows, D, _stream())
    _lib.check(st, "opb_l2_normalize_bwd")
    _count()
    return (dx16, dx32) if want_f32 else dx16


def text_embed_bwd(dx, tokens, dtable, dpos, dcls, pad_idx=1):
    B, T = tokens.shape
    D = dx.shape[-1]
    st = _lib.load().opb_text_embed_bwd(dx.data_ptr(), tokens.data_ptr(), dtable.data_ptr(), dpos.data_ptr(), dcls.data_ptr(),
                                        B, T, D, pad_idx, _stream())
    _lib.check(st, "opb_text_embed_bwd")
    _count()


def window_gather(x, B, t_in, t_out, stride, kw, pad, groups):
    """bf16 [B*t_in, C] -> [groups, B*t_out, kw*(C/groups)] convolution windows (see onepeace_b200.h)"""
    C = x.shape[1]
    cg = C // groups
    out = torch.empty(groups, B * t_out, kw * cg, dtype=torch.bfloat16, device=x.device)
    st = _lib.load().opb_window_gather(x.data_ptr(), out.data_ptr(), B, t_in, t_out, stride, kw, pad, groups, cg, _stream())
    _lib.check(st, "opb_window_gather")
    _count()
    return out


def window_scatter(dwin, B, t_in, t_out, stride, kw, pad):
    """adjoint of window_gather: bf16 [groups, B*t_out, kw*cg] -> [B*t_in, groups*cg]"""
    groups = dwin.shape[0]
    cg = dwin.shape[2] // kw
    dx = torch.empty(B * t_in, groups * cg, dtype=torch.bfloat16, device=dwin.device)
    st = _lib.load().opb_window_scatter(dwin.data_ptr(), dx.data_ptr(), B, t_in, t_out, stride, kw, pad, groups, cg, _stream())
    _lib.check(st, "opb_window_scatter")
    _count()
    return dx


def topk10_rows(sim, want_values=False):
    """fp32 [R, C] -> int32 [R, 10] column indices of the 10 largest entries per row (descending)"""
    _need_cuda(sim)
    assert sim.dtype == torch.float32 and sim.stride(1) == 1
    R, C = sim.shape
    idx = torch.empty(R, 10, dtype=torch.int32, device=sim.device)
    val = torch.empty(R, 10, dtype=torch.float32, device=sim.device) if want_values else None
    st = _lib.load().opb_topk10_rows(sim.data_ptr(), sim.stride(0), idx.data_ptr(), _ptr(val), R, C, _stream())
    _lib.check(st, "opb_topk10_rows")
    _count()
    return (idx, val) if want_values else idx


def recall_hits(idx, cand_ids, row_ids):
    """-> int32 [3]: rows whose id is among the ids of their top-1 / top-5 / top-10 candidates"""
    hits = torch.zeros(3, dtype=torch.int32, device=idx.device)
    st = _lib.load().opb_recall_hits(idx.data_ptr(), cand_ids.data_ptr(), row_ids.data_ptr(), idx.shape[0], hits.data_ptr(),
                                     _stream())
    _lib.check(st, "opb_recall_hits")
    _count()
    return hits


# ----------------------------------------------------------------------------------------------------------------
# pretraining path: row gathers, general dense relative-position bias
# ----------------------------------------------------------------------------------------------------------------
def row_gather(src, idx, out=None, fill=None, out_dtype=None, add=None):
    """out[r] = (src[idx[r]] if idx[r] >= 0 else fill (fp32 [dim]) / 0) + add[r % period] (add fp32 [period, dim] or None).
    src [n, dim] fp32 / bf16 (row stride free), idx int64 [rows]."""
    _need_cuda(src, idx)
    assert idx.dtype == torch.int64 and idx.is_contiguous() and src.stride(-1) == 1 and src.dim() == 2
    rows, dim = idx.numel(), src.shape[1]
    if out is None:
        out = torch.empty(rows, dim, dtype=out_dtype or src.dtype, device=src.device)
    if rows == 0:
        return out
    assert out.stride(-1) == 1 and (fill is None or (fill.dtype == torch.float32 and fill.is_contiguous()))
    assert add is None or (add.dtype == torch.float32 and add.is_contiguous() and add.shape[-1] == dim)
    st = _lib.load().opb_row_gather(src.data_ptr(), _dt(src), src.stride(0), idx.data_ptr(), _ptr(fill), _ptr(add),
                                    add.numel() // dim if add is not None else 0, out.data_ptr(), _dt(out), out.stride(-2),
                                    rows, dim, _stream())
    _lib.check(st, "opb_row_gather")
    _count()
    return out


def row_scatter_add(dout, idx, dsrc):
    """dsrc[idx[r]] += dout[r] for idx[r] >= 0; dsrc fp32 [n, dim] (caller zero-initialises)."""
    _need_cuda(dout, idx, dsrc)
    assert idx.dtype == torch.int64 and idx.is_contiguous() and dsrc.dtype == torch.float32 and dout.dim() == 2
    rows, dim = idx.numel(), dout.shape[1]
    if rows == 0:
        return dsrc
    st = _lib.load().opb_row_scatter_add(dout.data_ptr(), _dt(dout), dout.stride(0), idx.data_ptr(), dsrc.data_ptr(),
                                         dsrc.stride(-2), rows, dim, _stream())
    _lib.check(st, "opb_row_scatter_add")
    _count()
    return dsrc


def relpos_bias_block(table, bucket, ids, n, lo, bias, S, H):
    """Writes one modality's diagonal block of the dense bias canvas `bias` fp32 [Bb, H, S, s_pad] (Bb = 1 when ids is None):
    bias[bb, h, lo+i, lo+j] = table[bucket[p_i, p_j], h], p = ids[bb] (int64 [Bb, n], -1 = padded slot) or arange(n)."""
    _need_cuda(table, bucket, bias)
    assert table.dtype == torch.float32 and table.is_contiguous() and table.shape[1] == H and bias.is_contiguous()
    assert bucket.dtype == torch.int64 and bucket.stride(1) == 1 and bias.dim() == 4 and bias.shape[1] == H and bias.shape[2] == S
    Bb = bias.shape[0]
    if ids is not None:
        assert ids.dtype == torch.int64 and ids.shape == (Bb, n) and ids.stride(1) == 1
    st = _lib.load().opb_relpos_bias_block(table.data_ptr(), bucket.data_ptr(), bucket.stride(0), _ptr(ids),
                                           ids.stride(0) if ids is not None else 0, Bb, n, lo, bias.data_ptr(), S,
                                           bias.shape[3], H, _stream())
    _lib.check(st, "opb_relpos_bias_block")
    _count()
    return bias


def relpos_bias_block_bwd(dbias, bucket, ids, n, lo, dtable, S, H):
    Bb = dbias.shape[0]
    st = _lib.load().opb_relpos_bias_block_bwd(dbias.data_ptr(), bucket.data_ptr(), bucket.stride(0), _ptr(ids),
                                               ids.stride(0) if ids is not None else 0, Bb, n, lo, dtable.data_ptr(), S,
                                               dbias.shape[3], H, _stream())
    _lib.check(st, "opb_relpos_bias_block_bwd")
    _count()
    return dtable


def ln_fold(weight, ln_w, ln_b, bias, out_w, colsum, bias_out, interleave=0):
    """Folds LayerNorm(ln_w, ln_b) into the following Linear(weight, bias): writes the bf16 operand rows, their column sums and
    the fused bias into (views of) out_w / colsum / bias_out.  weight fp32 / bf16 [N, K]; ln_w / ln_b / bias fp32 or None."""
    _need_cuda(weight, out_w)
    N, Kd = weight.shape
    assert weight.stride(1) == 1 and out_w.dtype == torch.bfloat16 and out_w.stride(-1) == 1
    st = _lib.load().opb_ln_fold(weight.data_ptr(), _dt(weight), weight.stride(0), _ptr(ln_w), _ptr(ln_b), _ptr(bias), N, Kd,
                                 interleave, out_w.data_ptr(), out_w.stride(0), colsum.data_ptr(), bias_out.data_ptr(), _stream())
    _lib.check(st, "opb_ln_fold")
    _count()


def gemm_t(a, b, epi, out, a_mn=False, b_mn=False, bias=None, cta_group=0):
    """out[M, N] = A B^T with MN-major operands: a is [K, M] when a_mn else [M, K]; b is [K, N] when b_mn else [N, K]
    (bf16, unit column stride, free row pitch).  No transposed copies: TMA + UMMA read the row-contracted layout directly."""
    _need_cuda(a, b, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    Kd, M = (a.shape[0], a.shape[1]) if a_mn else (a.shape[1], a.shape[0])
    Kb, N = (b.shape[0], b.shape[1]) if b_mn else (b.shape[1], b.shape[0])
    assert Kd == Kb and out.shape == (M, N)
    st = _lib.load().opb_gemm_bf16_t(a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0), int(b_mn), M, N, Kd, epi,
                                     out.data_ptr(), out.stride(0), _ptr(bias), cta_group, _stream())
    _lib.check(st, "opb_gemm_bf16_t")
    _count()
    if PROFILE_HOOK is not None:
        PROFILE_HOOK("gemm", 2.0 * M * N * Kd, (M, N, Kd, epi))
    return out
