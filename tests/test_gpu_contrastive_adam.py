"""GPU parity of the contrastive head and the fused Adam step against (1) golden vectors produced by the
reference's own criterion / optimizer files and (2) oracle/restated.py at larger sizes."""
import math
import os

import pytest
import torch

import restated as R
import synth

pytestmark = pytest.mark.gpu


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")


def test_itc_loss_vs_reference_golden(golden_dir):
    """loss within 1e-3 relative (north_star), identical arg-max counts, gradients vs the reference's autograd."""
    need_gpu()
    from one_peace_b200.criterions.image_text_retrieval_loss import itc_loss
    cases = torch.load(os.path.join(golden_dir, "itc_loss.pt"), weights_only=False)
    for c in cases:
        a, t = synth.contrastive_pair(c["b"], c["d"], c["seed"])
        a = a.cuda().requires_grad_(True)
        t = t.cuda().requires_grad_(True)
        ls = c["logit_scale"].clone().cuda().requires_grad_(True)
        loss, i2t, t2i = itc_loss(a, t, a.detach(), t.detach(), ls.exp(), 0, c["eps"])
        loss.backward()
        rel = abs(loss.item() - c["loss"].item()) / abs(c["loss"].item())
        print(f"b={c['b']} d={c['d']} eps={c['eps']}: loss {loss.item():.6f} vs {c['loss'].item():.6f} rel {rel:.2e}")
        assert rel < 1e-3
        assert float(i2t) == float(c["i2t_ncorrect"]) and float(t2i) == float(c["t2i_ncorrect"])
        ga, gt = a.grad.cpu(), t.grad.cpu()
        # gradient factors pass through bf16 (2^-9 relative per element); compare in norm and direction
        assert abs(ga.norm().item() - c["grad_image_norm"].item()) / c["grad_image_norm"].item() < 5e-3
        assert abs(gt.norm().item() - c["grad_text_norm"].item()) / c["grad_text_norm"].item() < 5e-3
        assert torch.nn.functional.cosine_similarity(ga[:8].flatten(), c["grad_image"].flatten(), dim=0) > 0.9995
        assert torch.nn.functional.cosine_similarity(gt[:8].flatten(), c["grad_text"].flatten(), dim=0) > 0.9995
        assert abs(ls.grad.item() - c["grad_logit_scale"].item()) <= 5e-3 * abs(c["grad_logit_scale"].item()) + 1e-5


@pytest.mark.parametrize("b,world,rank,eps", [(256, 4, 2, 0.0), (1024, 8, 5, 0.0), (200, 2, 1, 0.1)])
def test_itc_loss_sharded_vs_oracle(b, world, rank, eps):
    """Config 4-i at full width: local b rows against W*b gathered rows (the gather is emulated by building all
    ranks' shards locally); targets are offset by rank*b."""
    need_gpu()
    from one_peace_b200.criterions.image_text_retrieval_loss import itc_loss
    d = 1536
    a_all, t_all = synth.contrastive_pair(b * world, d, seed=100 + b)
    a_loc = a_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    t_loc = t_all[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    ls = torch.tensor(math.log(1 / 0.07), requires_grad=True)
    want, wi, wt = R.itc_loss(a_loc, t_loc, a_all, t_all, R.logit_scale_exp(ls), rank, eps)
    want.backward()
    ga = a_loc.detach().cuda().requires_grad_(True)
    gt = t_loc.detach().cuda().requires_grad_(True)
    gls = ls.detach().cuda().requires_grad_(True)
    loss, i2t, t2i = itc_loss(ga, gt, a_all.cuda(), t_all.cuda(), gls.exp(), rank, eps)
    loss.backward()
    rel = abs(loss.item() - want.item()) / abs(want.item())
    print(f"b={b} W={world}: loss {loss.item():.6f} vs oracle {want.item():.6f} (rel {rel:.2e}); acc {float(i2t)}/{float(wi)}")
    assert rel < 1e-3
    # bf16 operands can flip an arg-max only on near-ties: allow none here (the synthetic pairs are well separated)
    assert float(i2t) == float(wi) and float(t2i) == float(wt)
    for got, ref in ((ga.grad.cpu(), a_loc.grad), (gt.grad.cpu(), t_loc.grad)):
        assert abs(got.norm() - ref.norm()) / ref.norm() < 5e-3
        assert torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0) > 0.9995
    assert abs(gls.grad.item() - ls.grad.item()) <= 5e-3 * abs(ls.grad.item()) + 1e-5


def test_criterion_forward_contract():
    """forward(model, sample) -> (loss, 1, logging dict with the reference's keys) with a stand-in model."""
    need_gpu()
    from one_peace_b200.criterions import ImageTextRetrievalCriterion
    a, t = synth.contrastive_pair(32, 256, seed=3)
    a, t = a.cuda(), t.cuda()

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.logit_scale = torch.nn.Parameter(torch.tensor(math.log(1 / 0.07), device="cuda"))
            self.a = torch.nn.Parameter(a.clone()); self.t = torch.nn.Parameter(t.clone())

        def forward(self, src_tokens=None, src_images=None, encoder_type=None, return_logit_scale=False):
            if return_logit_scale:
                return self.logit_scale.exp()
            return self.t if encoder_type == "text" else self.a
    m = M()
    crit = ImageTextRetrievalCriterion(task=None, label_smoothing=0.0)
    loss, ss, log = crit(m, {"net_input": {"src_tokens": None, "src_images": None}, "nsentences": 32})
    assert ss == 1 and set(log) == {"loss", "nsentences", "sample_size", "i2t_ncorrect", "t2i_ncorrect", "logit_scale_exp"}
    loss.backward()
    assert m.a.grad is not None and m.t.grad is not None and m.logit_scale.grad is not None
    want, _, _ = R.itc_loss(a.cpu(), t.cpu(), a.cpu(), t.cpu(), torch.tensor(1 / 0.07))
    assert abs(loss.item() - want.item()) / want.item() < 1e-3


@pytest.mark.parametrize("tag", ["fp32", "bf16"])
def test_adam_vs_reference_golden(golden_dir, tag):
    """Three steps of the reference python Adam (optim/adam.py executed as-is) vs the fused kernel."""
    need_gpu()
    from one_peace_b200.optim import Adam
    fx = torch.load(os.path.join(golden_dir, "adam.pt"), weights_only=False)[tag]
    p = torch.nn.Parameter(fx["p0"].clone().cuda())
    opt = Adam([p], lr=fx["lr"], betas=fx["betas"], eps=fx["eps"], weight_decay=fx["weight_decay"])
    for g, want in zip(fx["grads"], fx["traj"]):
        p.grad = g.clone().cuda()
        opt.step()
        if tag == "fp32":
            torch.testing.assert_close(p.detach().cpu(), want, atol=1e-7, rtol=1e-6)
        else:   # bf16 parameters: identical after rounding except for last-ulp ties
            diff = (p.detach().cpu().float() - want.float()).abs()
            assert (diff > 0).float().mean() < 0.01 and diff.max() <= want.float().abs().max() * 2 ** -7
    st = opt.state[p]
    torch.testing.assert_close(st["exp_avg"].cpu(), fx["exp_avg"], atol=1e-8, rtol=1e-6)
    torch.testing.assert_close(st["exp_avg_sq"].cpu(), fx["exp_avg_sq"], atol=1e-10, rtol=2e-6)


def test_adam_multi_tensor_groups_clip_and_master():
    """Many tensors (odd sizes, two groups with lr_scale / no-decay as utils/layer_decay.py builds them), grad-norm
    clipping folded into the step, fp32 master weights; oracle = restated.adam_step + clip_coefficient."""
    need_gpu()
    from one_peace_b200.optim import Adam, MemoryEfficientBF16Optimizer, AdjustAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(1536, 384), (1536,), (77,), (3, 5, 7), (8193,), (1,), (50000,)]
    ps = [torch.randn(s, generator=g) for s in shapes]
    gs = [[torch.randn(s, generator=g) * 0.3 for s in shapes] for _ in range(2)]

    class Cfg:
        lr = [1e-2]; adam_betas = "(0.9, 0.98)"; adam_eps = 1e-8; weight_decay = 0.05; master_weights = True
    params = [torch.nn.Parameter(p.clone().bfloat16().cuda()) for p in ps]
    groups = [dict(params=params[:3], weight_decay=0.05, lr_scale=0.5), dict(params=params[3:], weight_decay=0.0, lr_scale=1.0)]
    fo = AdjustAdam(Cfg, groups)
    fo.set_lr(1e-2)
    opt = MemoryEfficientBF16Optimizer(fo)
    # oracle state (fp32 master, python form)
    om = [p.bfloat16().float() for p in ps]
    m = [torch.zeros_like(p) for p in ps]; v = [torch.zeros_like(p) for p in ps]
    for step, grads in enumerate(gs, start=1):
        for p, gr in zip(params, grads):
            p.grad = gr.bfloat16().cuda()
        opt.multiply_grads(0.5)
        norm = opt.clip_grad_norm(1.0)
        opt.step()
        gb = [gr.bfloat16().float() for gr in grads]
        wnorm, coef = R.clip_coefficient(gb, 1.0, multiply_factor=0.5)
        assert abs(norm.item() - wnorm) / wnorm < 1e-5
        for i in range(len(ps)):
            lr = 1e-2 * (0.5 if i < 3 else 1.0)
            wd = 0.05 if i < 3 else 0.0
            R.adam_step(om[i], gb[i] * (0.5 * coef), m[i], v[i], step, lr, 0.9, 0.98, 1e-8, wd)
    for i, p in enumerate(params):
        master = opt.optimizer.state[p]["master"].cpu()
        torch.testing.assert_close(master, om[i], atol=2e-6, rtol=2e-5)
        assert torch.equal(p.detach().cpu(), master.bfloat16())
    # determinism of the norm: same grads -> bit-identical norm
    n1 = opt.optimizer.grad_norm_and_scale(1.0, 0.0)[0].item()
    n2 = opt.optimizer.grad_norm_and_scale(1.0, 0.0)[0].item()
    assert n1 == n2


def test_two_rank_contrastive_step_nccl():
    """configs[3] at W = 2 on real GPUs: NCCL all-gather + InfoNCE fwd/bwd vs the oracle (skipped on a 1-GPU box;
    profiles/r01_dist_contrastive_2gpu.log holds the round-1 run)."""
    need_gpu()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(root, "scripts", "dist_contrastive.py"), "--b", "256",
                        "--steps", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok=True" in r.stdout


def test_two_rank_sharded_adam_nccl():
    """optim/distributed_adam.py at W = 2 on real GPUs: reduce-scatter + fused shard step + all-gather == un-sharded fused
    Adam on the averaged gradients, incl. global-norm clipping (skipped on a 1-GPU box)."""
    need_gpu()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", os.path.join(root, "scripts", "dist_zero_adam.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok=True" in r.stdout


def test_forward_after_optimizer_steps_uses_updated_weights():
    """the fused Adam kernel writes parameters through raw pointers; every cached kernel-ready pack
    (concatenated QKV, LN-folded weights, fp32 LayerNorm copies, rel-pos LUTs) must be rebuilt.  Two steps, then the
    inference forward and the training forward must both equal the oracle evaluated on the optimizer's own parameters."""
    need_gpu()
    import synth
    from one_peace_b200.one_peace.hub_interface import from_pretrained
    from one_peace_b200.optim import Adam
    cfgd = dict(embed_dim=256, ffn=1024, layers=2, heads=4)
    sd = synth.make_state_dict(**cfgd, modalities=("text",), seed=3)
    hub = from_pretrained(state_dict=sd, head_type="text", layers=2, embed_dim=256, ffn_embed_dim=1024, attention_heads=4,
                          device="cuda", dtype="float32")
    model = hub.model
    tok, _, _, _ = synth.tiny_inputs(seed=0, n_text=8)
    tok = tok.cuda()
    opt = Adam(model.parameters(), lr=3e-2, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.0)
    g = torch.Generator().manual_seed(5)
    target = torch.randn(8, 256, generator=g).cuda()
    with torch.no_grad():
        before = model(src_tokens=tok, encoder_type="text").float().clone()
    model.train()
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        (model(src_tokens=tok, encoder_type="text").float() * target).sum().backward()
        opt.step()
    model.eval()
    with torch.no_grad():
        got = model(src_tokens=tok, encoder_type="text").float().cpu()
    model.train()
    got_train = model(src_tokens=tok, encoder_type="text").float().detach().cpu()
    assert torch.nn.functional.cosine_similarity(got, before.cpu()).min() < 0.999, "lr 3e-2 x 2 steps must move the embeddings"
    cfg = R.OracleConfig(embed_dim=256, ffn_embed_dim=1024, layers=2, attention_heads=4)
    sd_now = {k: (v.detach().float() if v.is_floating_point() else v.detach()).cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        want = R.extract_features(sd_now, cfg, "text", src_tokens=tok.cpu())
    assert torch.nn.functional.cosine_similarity(got, want).min() > 0.999
    assert torch.nn.functional.cosine_similarity(got_train, want).min() > 0.999


def test_adam_state_survives_load_state_dict_with_bf16_params():
    """torch's Optimizer.load_state_dict casts state to the parameter dtype; exp_avg / exp_avg_sq / the
    fp32 master must come back as the saved fp32 tensors (reference: fp16_optimizer_memory_efficent.py:44-62)."""
    need_gpu()
    from one_peace_b200.optim import Adam
    g = torch.Generator().manual_seed(1)
    mk = lambda: [torch.nn.Parameter(torch.randn(1000, generator=torch.Generator().manual_seed(2)).bfloat16().cuda()),
                  torch.nn.Parameter(torch.randn(37, 5, generator=torch.Generator().manual_seed(3)).bfloat16().cuda())]
    grads = [[torch.randn(p.shape, generator=g).bfloat16().cuda() for p in mk()] for _ in range(3)]
    pa = mk()
    oa = Adam(pa, lr=1e-2, betas=(0.9, 0.98), weight_decay=0.05, master_weights=True)
    for p, gr in zip(pa, grads[0]):
        p.grad = gr.clone()
    oa.step()
    state = oa.state_dict()
    state = {"state": {k: {n: (t.clone() if torch.is_tensor(t) else t) for n, t in v.items()} for k, v in state["state"].items()},
             "param_groups": state["param_groups"]}
    pb = mk()
    with torch.no_grad():
        for q, p in zip(pb, pa):
            q.copy_(p)
    ob = Adam(pb, lr=1e-2, betas=(0.9, 0.98), weight_decay=0.05, master_weights=True)
    ob.load_state_dict(state)
    for q, p in zip(pb, pa):
        for n in ("exp_avg", "exp_avg_sq", "master"):
            assert ob.state[q][n].dtype == torch.float32 and torch.equal(ob.state[q][n], oa.state[p][n]), n
        assert ob.state[q]["step"] == 1
    for step in (1, 2):
        for p, q, gr in zip(pa, pb, grads[step]):
            p.grad = gr.clone(); q.grad = gr.clone()
        oa.step(); ob.step()
    for p, q in zip(pa, pb):
        assert torch.equal(p.detach(), q.detach()) and torch.equal(oa.state[p]["master"], ob.state[q]["master"])


def test_adam_per_parameter_step_counts():
    """a parameter that gets its first gradient later than its group-mates keeps its own bias correction
    (optim/adam.py:207-213 tracks `step` per parameter)."""
    need_gpu()
    from one_peace_b200.optim import Adam
    g = torch.Generator().manual_seed(4)
    a0, b0 = torch.randn(300, generator=g), torch.randn(200, generator=g)
    ga = [torch.randn(300, generator=g) for _ in range(3)]
    gb = [None, torch.randn(200, generator=g), torch.randn(200, generator=g)]
    pa, pb = torch.nn.Parameter(a0.clone().cuda()), torch.nn.Parameter(b0.clone().cuda())
    opt = Adam([pa, pb], lr=1e-2, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
    wa, wb = a0.clone(), b0.clone()
    ma, va, mb, vb = torch.zeros(300), torch.zeros(300), torch.zeros(200), torch.zeros(200)
    tb = 0
    for t in range(3):
        pa.grad = ga[t].cuda()
        pb.grad = None if gb[t] is None else gb[t].cuda()
        opt.step()
        R.adam_step(wa, ga[t], ma, va, t + 1, 1e-2, 0.9, 0.98, 1e-8, 0.05)
        if gb[t] is not None:
            tb += 1
            R.adam_step(wb, gb[t], mb, vb, tb, 1e-2, 0.9, 0.98, 1e-8, 0.05)
    torch.testing.assert_close(pa.detach().cpu(), wa, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(pb.detach().cpu(), wb, atol=1e-6, rtol=1e-5)
    assert opt.state[pa]["step"] == 3 and opt.state[pb]["step"] == 2
