"""GPU: the hand-written adjoint kernels (csrc/backward.cu, csrc/attention_bwd.cu) through the C-ABI vs torch autograd
of the same op in fp32.  Tolerances: inputs that are bf16 in the product path are rounded to bf16 BEFORE the fp32
reference runs, so the only differences are bf16 output rounding (2^-9) and fp32 summation order."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from one_peace_b200 import kernels
    return kernels


def relerr(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-9)).item()


def gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("rows,dim,xdt,dydt,dxdt,gelu", [
    (300, 1536, torch.float32, torch.bfloat16, torch.float32, False),
    (1000, 6144, torch.bfloat16, torch.bfloat16, torch.bfloat16, False),
    (777, 384, torch.bfloat16, torch.bfloat16, torch.bfloat16, True),
    (5, 256, torch.float32, torch.float32, torch.float32, False),
])
def test_layernorm_bwd(K, rows, dim, xdt, dydt, dxdt, gelu):
    g = gen(rows + dim)
    x = (torch.randn(rows, dim, device="cuda", generator=g) * 1.3 + 0.2).to(xdt)
    dy = torch.randn(rows, dim, device="cuda", generator=g).to(dydt)
    w = 1 + 0.2 * torch.randn(dim, device="cuda", generator=g)
    b = 0.1 * torch.randn(dim, device="cuda", generator=g)
    xr = x.float().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xr, (dim,), wr, br, 1e-5)
    if gelu:
        y = torch.nn.functional.gelu(y)
    y.backward(dy.float())
    dx = torch.empty(rows, dim, device="cuda", dtype=dxdt)
    dg = torch.empty(dim, device="cuda"); db = torch.empty(dim, device="cuda")
    K.layernorm_bwd(x, dy, w, b, dx, eps=1e-5, gelu=gelu, dgamma=dg, dbeta=db)
    tol = 6e-3 if dxdt == torch.bfloat16 else 2e-5
    assert relerr(dx, xr.grad) < tol
    assert relerr(dg, wr.grad) < 1e-4 and relerr(db, br.grad) < 1e-4
    if dxdt == torch.float32:       # accumulate into an existing fp32 gradient, no parameter grads requested
        acc = torch.ones(rows, dim, device="cuda")
        K.layernorm_bwd(x, dy, w, b, acc, eps=1e-5, gelu=gelu, accumulate=True)
        assert relerr(acc, 1 + xr.grad) < 2e-5


def test_geglu_fwd_bwd(K):
    rows, F = 333, 1024
    g = gen(7)
    gl = (torch.randn(rows, 2 * F, device="cuda", generator=g) * 1.5).bfloat16()
    du = torch.randn(rows, F, device="cuda", generator=g).bfloat16()
    glr = gl.float().requires_grad_(True)
    u_ref = torch.nn.functional.gelu(glr[:, :F]) * glr[:, F:]
    u_ref.backward(du.float())
    u = torch.empty(rows, F, device="cuda", dtype=torch.bfloat16)
    dgl = torch.empty(rows, 2 * F, device="cuda", dtype=torch.bfloat16)
    K.geglu_fwd(gl, u)
    K.geglu_bwd(gl, du, dgl)
    assert relerr(u, u_ref) < 6e-3
    assert relerr(dgl, glr.grad) < 6e-3


@pytest.mark.parametrize("with_scale", [False, True])
def test_scale_resid_fwd_bwd(K, with_scale):
    rows, n = 1234, 1536
    g = gen(11)
    x = torch.randn(rows, n, device="cuda", generator=g)
    o = torch.randn(rows, n, device="cuda", generator=g).bfloat16()
    gamma = torch.randn(n, device="cuda", generator=g)
    rs = None
    if with_scale:    # drop-path: per-row keep mask / keep_prob (transformer_layer.py:80-86)
        rs = (torch.rand(rows, device="cuda", generator=g) < 0.6).float() / 0.6
    dx = torch.randn(rows, n, device="cuda", generator=g)
    out = torch.empty_like(x)
    K.scale_resid_fwd(x, o, gamma, rs, out)
    scale = rs[:, None] if with_scale else 1.0
    want = x + scale * gamma * o.float()
    assert relerr(out, want) < 1e-6
    d_o = torch.empty(rows, n, device="cuda", dtype=torch.bfloat16)
    dgamma = torch.empty(n, device="cuda"); dbias = torch.empty(n, device="cuda")
    K.scale_resid_bwd(dx, o, gamma, rs, d_o, dgamma=dgamma, dbias=dbias)
    do_ref = scale * gamma * dx
    assert relerr(d_o, do_ref) < 6e-3
    assert relerr(dgamma, (scale * dx * o.float()).sum(0)) < 1e-4
    assert relerr(dbias, do_ref.sum(0)) < 1e-4
    cs = torch.empty(n, device="cuda")
    K.colsum(d_o, cs)
    assert relerr(cs, d_o.float().sum(0)) < 1e-5


def attention_ref(qkv, bias, key_pad, B, S, H):
    """fp32 reference of multihead_attention.py:107-115 on the stored (already scaled) q."""
    D = H * 64
    q, k, v = [t.view(B, S, H, 64).transpose(1, 2) for t in qkv.view(B, S, 3 * D).split(D, dim=-1)]
    s = q @ k.transpose(-1, -2)
    if bias is not None:
        s = s + bias[None, :, :, :S]
    if key_pad is not None:
        s = s.masked_fill(key_pad.bool()[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B * S, D)


@pytest.mark.parametrize("B,S,H,pad,use_bias", [(3, 197, 4, False, True), (2, 70, 2, True, True), (2, 300, 2, True, False),
                                               (1, 64, 1, False, True)])
def test_attention_bwd(K, B, S, H, pad, use_bias):
    D = H * 64
    g = gen(S + H)
    qkv = (torch.randn(B * S, 3 * D, device="cuda", generator=g) * 0.6).bfloat16()
    s_pad = (S + 3) // 4 * 4
    bias = (torch.randn(H, S, s_pad, device="cuda", generator=g) * 0.5) if use_bias else None
    key_pad = None
    if pad:
        key_pad = torch.zeros(B, S, dtype=torch.uint8, device="cuda")
        key_pad[0, S - 5:] = 1
        key_pad[-1, S // 2:] = 1
    d_out = torch.randn(B * S, D, device="cuda", generator=g).bfloat16()
    lse = torch.empty(B * H * S, device="cuda")
    out = K.attention(qkv, bias, key_pad, B, S, H, lse=lse)
    # reference
    qr = qkv.float().requires_grad_(True)
    br = bias.clone().requires_grad_(True) if use_bias else None
    o_ref = attention_ref(qr, br, key_pad, B, S, H)
    assert relerr(out, o_ref) < 1e-2
    o_ref.backward(d_out.float())
    q_scale = 0.125
    dqkv = torch.zeros(B * S, 3 * D, device="cuda", dtype=torch.bfloat16)
    dbias = torch.zeros(H, S, s_pad, device="cuda") if use_bias else None
    K.attention_bwd(qkv, out, d_out, bias, key_pad, lse, dqkv, dbias, B, S, H, q_scale)
    want = qr.grad.clone()
    want[:, :D] *= q_scale
    # P and dS pass through bf16 on the tensor cores (2^-9 relative each); gradients are sums of ~S such terms
    for name, lo in (("dq", 0), ("dk", D), ("dv", 2 * D)):
        err = relerr(dqkv[:, lo:lo + D], want[:, lo:lo + D])
        assert err < 1.5e-2, (name, err)
    if use_bias:
        # fp32 atomics of fp32 dS; delta = sum(dO * O) uses the bf16-rounded forward output (2^-9 per element)
        assert relerr(dbias[:, :, :S], br.grad[:, :, :S]) < 1e-2
        if s_pad > S:
            assert torch.all(dbias[:, :, S:] == 0)


@pytest.mark.parametrize("B,S,H,pad", [(3, 197, 4, False), (2, 70, 2, True), (2, 214, 3, True), (5, 33, 2, True)])
def test_attention_bwd_transposed_tables(K, B, S, H, pad):
    """opb_attention_bwd_t (tcgen05, bias / dbias as transposed tables shared by several launches) vs fp32 torch autograd; the
    gradient table is accumulated over TWO launches and folded back once, as the encoder stack does."""
    D = H * 64
    g = gen(S + H + 1)
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
    qr = qkv.float().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    attention_ref(qr, br, key_pad, B, S, H).backward(d_out.float())
    bias_t = K.relpos_bias_transpose(bias)
    dbias_t = torch.zeros(H, K.BIAS_T_KEYS, K.BIAS_T_Q, device="cuda")
    dqkv = torch.zeros(B * S, 3 * D, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        K.attention_bwd_t(qkv, out, d_out, bias_t, key_pad, lse, dqkv, dbias_t, B, S, H, 0.125)
    dbias = torch.zeros(H, S, s_pad, device="cuda")
    K.relpos_dbias_fold(dbias_t, dbias)
    want = qr.grad.clone()
    want[:, :D] *= 0.125
    for name, lo in (("dq", 0), ("dk", D), ("dv", 2 * D)):
        err = relerr(dqkv[:, lo:lo + D], want[:, lo:lo + D])
        assert err < 1.5e-2, (name, err)
    assert relerr(dbias[:, :, :S], 2 * br.grad[:, :, :S]) < 1e-2
    if s_pad > S:
        assert torch.all(dbias[:, :, S:] == 0)
    # everything outside the S x S corner of the transposed table is padding and must stay zero
    assert torch.all(dbias_t[:, S:, :] == 0) and torch.all(dbias_t[:, :, S:] == 0)


def test_relpos_dbias_center(K):
    """zero-row-sum projection of the accumulated bias gradient: every (head, query) row loses its mean over the S valid columns,
    the padding columns stay untouched"""
    H, S, s_pad = 3, 197, 200
    g = gen(21)
    db = torch.randn(H, S, s_pad, device="cuda", generator=g)
    db[:, :, S:] = 7.0
    want = db.clone()
    want[:, :, :S] -= want[:, :, :S].mean(-1, keepdim=True)
    K.relpos_dbias_center(db)
    torch.testing.assert_close(db, want, atol=1e-5, rtol=1e-5)
    assert db[:, :, :S].sum(-1).abs().max() < 1e-3


def test_relpos_bias_bwd(K):
    S, H, NB = 50, 4, 37
    g = gen(3)
    bucket = torch.randint(0, NB, (64, 64), device="cuda", generator=g)
    s_pad = 52
    dbias = torch.randn(H, S, s_pad, device="cuda", generator=g)
    dtable = torch.zeros(NB, H, device="cuda")
    K.relpos_bias_bwd(dbias, bucket, dtable, S)
    want = torch.zeros(NB, H, device="cuda")
    want.index_add_(0, bucket[:S, :S].reshape(-1), dbias[:, :, :S].permute(1, 2, 0).reshape(S * S, H))
    assert relerr(dtable, want) < 1e-5


@pytest.mark.parametrize("B,t_in,stride,kw,pad,groups,cg", [(2, 49, 1, 19, 9, 16, 16), (3, 99, 2, 3, 0, 1, 64), (2, 24, 2, 2, 0, 1, 512)])
def test_window_gather_scatter(K, B, t_in, stride, kw, pad, groups, cg):
    """im2col / col2im of the audio adapter's training path vs an index-built reference and its autograd adjoint."""
    t_out = (t_in + 2 * pad - kw) // stride + 1
    C = groups * cg
    g = gen(t_in + kw)
    x = torch.randn(B * t_in, C, device="cuda", generator=g).bfloat16()
    win = K.window_gather(x, B, t_in, t_out, stride, kw, pad, groups)
    xr = x.float().view(B, t_in, groups, cg).requires_grad_(True)
    xp = torch.nn.functional.pad(xr, (0, 0, 0, 0, pad, pad))                       # pad the time axis
    idx = (torch.arange(t_out, device="cuda")[:, None] * stride + torch.arange(kw, device="cuda")[None, :])   # [t_out, kw]
    ref = xp[:, idx]                                                                # [B, t_out, kw, groups, cg]
    ref = ref.permute(3, 0, 1, 2, 4).reshape(groups, B * t_out, kw * cg)
    assert torch.equal(win.float(), ref.detach())
    dwin = torch.randn(groups, B * t_out, kw * cg, device="cuda", generator=g).bfloat16()
    ref.backward(dwin.float())
    dx = K.window_scatter(dwin, B, t_in, t_out, stride, kw, pad)
    assert relerr(dx, xr.grad.reshape(B * t_in, C)) < 6e-3
