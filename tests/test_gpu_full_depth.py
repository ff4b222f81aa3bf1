"""GPU: BASELINE.json configs[1] and [2] at FULL depth (40 layers at the 4B width, 4 distinct seeded layers cycled — drawing
3.9 B parameters takes minutes) against the fp32 CPU oracle, with the gates `north_star` states: cosine >= 0.999 per
modality, InfoNCE loss within 1e-3 relative, retrieval arg-max identical on EVERY row.

Two synthetic networks (scripts/parity_depth.py):
  * "conditioned" — LayerScale in (1e-3, 3e-3): per-layer updates small against the residual stream, the regime of a trained
    ONE-PEACE (LayerScale is initialised at 1e-6, finetune_3B.yaml:132).  The gates are asserted here.  Rounding the WEIGHTS
    to bf16 inside the fp32 oracle alone moves its text embeddings by 1 - cos = 2e-4 on this network.
  * "hard" — LayerScale U(0.5, 1.5): forty random O(1) residual branches amplify a 2^-9 perturbation ~20x (the fp32 oracle
    with bf16-rounded weights is itself only at cosine 0.987 to the fp32 oracle), so no bf16 implementation can meet 0.999;
    asserted instead: the error-budget control — the sm_100a path must be at least as close to the fp32 oracle as the
    reference's own arithmetic run in bf16 eager on the same GPU (oracle/restated.py on CUDA, bf16 weights + activations).
Arg-max on every row: candidates are drawn until every row (and column) of the oracle's similarity matrix is decided by a
margin > 4e-3 (>= 4 sigma of the bf16 path's similarity error), then ALL rows are compared (no row is skipped)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

import restated as R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import parity_depth as PD  # noqa: E402

pytestmark = pytest.mark.gpu


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")


def hub_for(sd, dtype="bfloat16", head_type="val"):
    from one_peace_b200.one_peace.hub_interface import from_pretrained
    return from_pretrained(state_dict=sd, head_type=head_type, layers=PD.L, embed_dim=PD.D, ffn_embed_dim=PD.FFN,
                           attention_heads=PD.H, patch_image_size=224, device="cuda", dtype=dtype, vocab_size=PD.VOCAB)


def pick_decided(sim, k, margin, both_ways=True, tries=40000, seed=0):
    """Candidate columns of the ORACLE similarity matrix (rows = queries) such that every row's best match — and, with
    both_ways, every kept column's best row — wins by more than `margin`: best of `tries` random k-subsets (batched), k
    reduced by two if no subset reaches the margin.  The construction is asserted, then EVERY row / column is compared."""
    g = torch.Generator().manual_seed(seed)
    n = sim.shape[1]
    while k >= 2:
        perm = torch.rand(tries, n, generator=g).argsort(dim=1)[:, :k]
        sub = sim[:, perm].permute(1, 0, 2)                                  # tries x rows x k
        tr = sub.topk(2, dim=2).values
        ok = (tr[..., 0] - tr[..., 1]).min(dim=1).values
        if both_ways:
            tc = sub.topk(2, dim=1).values
            ok = torch.minimum(ok, (tc[:, 0] - tc[:, 1]).min(dim=1).values)
        best = int(ok.argmax())
        if float(ok[best]) > margin:
            return perm[best].sort().values
        k -= 2
    raise AssertionError("no decided candidate set found")


def test_config3_trimodal_40_layers_gates_vs_fp32_oracle():
    need_gpu()
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    sd = PD.build_sd(PD.NETS["conditioned"])
    tok, img, aud, apm = PD.inputs(n_text=48, audio=True)
    want = PD.oracle_embeddings(sd, tok, img, aud, apm)                       # fp32, CPU
    hub = hub_for(sd)
    got = {"text": hub.extract_text_features(tok.cuda()).float().cpu(),
           "image": hub.extract_image_features(img.cuda()).float().cpu(),
           "audio": hub.extract_audio_features(aud.cuda(), apm.cuda()).float().cpu()}
    for m in ("text", "image", "audio"):
        cos = F.cosine_similarity(got[m], want[m]).min().item()
        assert cos >= 0.999, (m, cos)
    scale = R.logit_scale_exp(sd["logit_scale"])
    from one_peace_b200.criterions.image_text_retrieval_loss import itc_loss
    for a in ("image", "audio"):
        cols = pick_decided(want[a] @ want["text"].t(), PD.B, 4e-3)
        ws, gs = want[a] @ want["text"][cols].t(), got[a] @ got["text"][cols].t()
        assert torch.equal(gs.argmax(1), ws.argmax(1)) and torch.equal(gs.argmax(0), ws.argmax(0)), a      # every row / column
        wt, gt = want["text"][:PD.B], got["text"][:PD.B]                      # the config's 8 pairs for the loss
        lw = R.itc_loss(want[a], wt, want[a], wt, scale, 0, 0.0)[0].item()
        lg = itc_loss(got[a].cuda(), gt.cuda(), got[a].cuda(), gt.cuda(), scale.cuda(), 0, 0.0)[0].item()
        assert abs(lg - lw) / abs(lw) <= 1e-3, (a, lg, lw)


def test_config2_vision_batch64_40_layers_vs_fp32_oracle():
    """The benchmarked configuration (64 x 224^2 images, 4B vision branch, bf16): every image embedding of the batch-64 forward
    equals the fp32 oracle's (computed for the first 16 images: the encoder is per-sample independent) to cosine >= 0.999,
    identical nearest-neighbour arg-max on every query row, and a sample's embedding does not depend on the batch size."""
    need_gpu()
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    sd = PD.build_sd(PD.NETS["conditioned"], modalities=("image",))
    g = torch.Generator().manual_seed(0)
    img = torch.randn(64, 3, 224, 224, generator=g)
    hub = hub_for(sd, head_type="image")
    got = hub.extract_image_features(img.cuda()).float().cpu()
    cfg = R.OracleConfig(embed_dim=PD.D, ffn_embed_dim=PD.FFN, layers=PD.L, attention_heads=PD.H)
    with torch.no_grad():
        want = R.extract_features(sd, cfg, "image", src_images=img[:16])
    cos = F.cosine_similarity(got[:16], want).min().item()
    assert cos >= 0.999, cos
    # image-to-image retrieval inside the first 16 (self-match removed), gallery drawn until every query is decided by > 5e-3
    ws, gs = want @ want.t() - 2 * torch.eye(16), got[:16] @ got[:16].t() - 2 * torch.eye(16)
    cols = pick_decided(ws, 8, 4e-3, both_ways=False)
    assert torch.equal(gs[:, cols].argmax(1), ws[:, cols].argmax(1))
    small = hub.extract_image_features(img[:16].cuda()).float().cpu()
    assert F.cosine_similarity(small, got[:16]).min() > 0.99999           # batch size does not change a sample's embedding


def test_hard_network_error_budget_vs_eager_bf16():
    """Error-budget control on the ill-conditioned network: the sm_100a path (bf16 weights, bf16 GEMM operands, fp32 residual
    stream) must not be further from the fp32 oracle than the reference's arithmetic in bf16 eager on the same GPU."""
    need_gpu()
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    sd = PD.build_sd(PD.NETS["hard"], modalities=("text", "image"))
    tok, img, _, _ = PD.inputs()
    want = PD.oracle_embeddings(sd, tok, img, None, None)
    eager = PD.oracle_embeddings(sd, tok, img, None, None, "cuda", torch.bfloat16)
    hub = hub_for(sd, head_type="vl")
    got = {"text": hub.extract_text_features(tok.cuda()).float().cpu(), "image": hub.extract_image_features(img.cuda()).float().cpu()}
    for m in ("text", "image"):
        err_repo = 1.0 - F.cosine_similarity(got[m], want[m]).min().item()
        err_eager = 1.0 - F.cosine_similarity(eager[m], want[m]).min().item()
        assert err_repo <= err_eager + 1e-4, (m, err_repo, err_eager)
