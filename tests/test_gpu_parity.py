"""GPU parity of the public path (extract_*_features through the sm_100a kernels) against
 (1) tests/golden/tiny_retrieval.pt — outputs of the reference's own module files, and
 (2) oracle/restated.py on the same seeded inputs, incl. a 4B-width (d=1536, h=24, ffn=6144) slice.
Bars: cosine >= 0.999 per embedding, identical retrieval arg-max; max-abs diff reported."""
import os

import pytest
import torch

import restated as R
import synth

pytestmark = pytest.mark.gpu


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")


def build_hub(sd, head_type, layers, d, ffn, heads, dtype="float32"):
    from one_peace_b200.one_peace.hub_interface import from_pretrained
    return from_pretrained(state_dict=sd, head_type=head_type, layers=layers, embed_dim=d, ffn_embed_dim=ffn,
                           attention_heads=heads, patch_image_size=224, device="cuda", dtype=dtype)


def check(got, want, what, min_cos=0.999):
    got = got.float().cpu()
    cos = torch.nn.functional.cosine_similarity(got, want).min().item()
    mad = (got - want).abs().max().item()
    print(f"{what}: min cosine {cos:.6f}, max abs diff {mad:.3e}")
    assert cos >= min_cos, (what, cos)
    torch.testing.assert_close(got.norm(dim=1), torch.ones(got.shape[0]), atol=2e-3, rtol=0)
    return cos


@pytest.fixture(scope="module")
def tiny(golden_dir):
    need_gpu()
    fx = torch.load(os.path.join(golden_dir, "tiny_retrieval.pt"), weights_only=False)
    # same generator stream as the golden run (all three modalities); the 'vl' model drops the audio keys
    sd = synth.make_state_dict(**fx["config"], seed=fx["weights_seed"])
    hub = build_hub(sd, "vl", 2, 256, 1024, 4)
    return fx, sd, hub, synth.tiny_inputs(seed=fx["inputs_seed"])


def test_tiny_text_features_vs_reference_golden(tiny):
    fx, sd, hub, (tok, img, aud, apm) = tiny
    got = hub.extract_text_features(tok)
    check(got, fx["outputs"]["text"], "tiny text vs reference golden")


def test_tiny_image_features_vs_reference_golden(tiny):
    fx, sd, hub, (tok, img, aud, apm) = tiny
    got = hub.extract_image_features(img)
    check(got, fx["outputs"]["image"], "tiny image vs reference golden")


def test_tiny_adapter_outputs(tiny):
    fx, sd, hub, (tok, img, aud, apm) = tiny
    ew = hub.model.encoder_wrapper
    x, pad, bias = ew.text_adapter(tok.cuda())
    want = fx["adapter"]["text_x"] * (~fx["adapter"]["text_pad"]).unsqueeze(-1)
    torch.testing.assert_close(x.cpu(), want, atol=1e-6, rtol=0)
    S = x.shape[1]
    torch.testing.assert_close(bias[0][:, :, :S].cpu(), fx["adapter"]["text_bias"], atol=0, rtol=0)
    xi, _, bi = ew.image_adapter(img.cuda())
    ref = fx["adapter"]["image_x"]
    err = (xi[:1].cpu() - ref).abs().max().item() / ref.abs().max().item()
    print("image adapter rel err", err)
    assert err < 2e-2       # three bf16 GEMMs + two LN/GELU stages
    torch.testing.assert_close(bi[0][:, :40, :40].cpu(), fx["adapter"]["image_bias"], atol=0, rtol=0)


def test_tiny_retrieval_argmax_matches_oracle(tiny):
    fx, sd, hub, (tok, img, aud, apm) = tiny
    g = torch.Generator().manual_seed(9)
    imgs = torch.randn(8, 3, 224, 224, generator=g)
    cfg = R.OracleConfig(embed_dim=256, ffn_embed_dim=1024, layers=2, attention_heads=4)
    with torch.no_grad():
        wt = R.extract_features(sd, cfg, "text", src_tokens=tok)
        wi = R.extract_features(sd, cfg, "image", src_images=imgs)
    gt = hub.extract_text_features(tok).float().cpu()
    gi = hub.extract_image_features(imgs).float().cpu()
    check(gt, wt, "text vs oracle"); check(gi, wi, "image vs oracle")
    assert torch.equal((gt @ gi.t()).argmax(1), (wt @ wi.t()).argmax(1))
    assert torch.equal((gi @ gt.t()).argmax(1), (wi @ wt.t()).argmax(1))


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_4b_width_slice_vs_oracle(dtype):
    """d=1536 / 24 heads / ffn=6144 (the 4B layer shape), 3 layers, 4 images + 6 token sequences."""
    need_gpu()
    sd = synth.make_state_dict(embed_dim=1536, ffn=6144, layers=3, heads=24, modalities=("text", "image"), seed=4,
                               vocab=2048)
    hub = build_hub(sd, "vl", 3, 1536, 6144, 24, dtype=dtype)
    cfg = R.OracleConfig(embed_dim=1536, ffn_embed_dim=6144, layers=3, attention_heads=24)
    tok, img, _, _ = synth.tiny_inputs(seed=2, n_text=6, text_len=20, n_img=4, vocab=2048)
    with torch.no_grad():
        if dtype == "bfloat16":      # the oracle sees the same (bf16-rounded) parameters
            sdo = {k: (v.bfloat16().float() if v.is_floating_point() else v) for k, v in sd.items()}
        else:
            sdo = sd
        wt = R.extract_features(sdo, cfg, "text", src_tokens=tok)
        wi = R.extract_features(sdo, cfg, "image", src_images=img)
    gt = hub.extract_text_features(tok)
    gi = hub.extract_image_features(img)
    check(gt, wt, f"4B-width text ({dtype})"); check(gi, wi, f"4B-width image ({dtype})")
    gt, gi = gt.float().cpu(), gi.float().cpu()
    assert torch.equal((gt @ gi.t()).argmax(1), (wt @ wi.t()).argmax(1))


def test_forward_refuses_to_build_a_fake_graph(tiny):
    fx, sd, hub, (tok, img, aud, apm) = tiny
    with pytest.raises(NotImplementedError):
        hub.model(src_tokens=tok.cuda(), encoder_type="text")        # grad enabled: backward not built yet
