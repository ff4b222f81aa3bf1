"""CPU: the drop-in contract of the model class that does not need a GPU — parameter names and shapes
equal the reference's (checkpoint compatibility, checked against the names recorded from the reference's own modules in
tests/golden/tiny_train_grads.pt), state-dict upgrade / pruning per head type, and the absence of any CPU compute path."""
import os

import pytest
import torch

import synth

TINY = dict(embed_dim=256, ffn=1024, layers=2, heads=4)


def build(head_type, sd=None, device="cpu"):
    from one_peace_b200.one_peace.hub_interface import from_pretrained
    return from_pretrained(state_dict=sd, head_type=head_type, layers=2, embed_dim=256, ffn_embed_dim=1024,
                           attention_heads=4, patch_image_size=224, device=device).model


def test_parameter_names_and_shapes_match_the_reference(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "tiny_train_grads.pt"), weights_only=False)
    sd = synth.make_state_dict(**TINY, seed=0)
    model = build("val", sd)
    own = {n: tuple(p.shape) for n, p in model.named_parameters()}
    for modality in ("text", "image"):
        for name, summ in fx["grads"][modality].items():            # names / shapes recorded from the reference's modules
            assert name in own, name
            assert own[name] == tuple(summ["shape"]), (name, own[name], summ["shape"])
    # every tensor of the synthetic reference-layout state dict was consumed (strict load) and round-trips
    got = model.state_dict()
    for k, v in sd.items():
        assert k in got and tuple(got[k].shape) == tuple(v.shape), k
        assert torch.equal(got[k].float().cpu(), v.float()), k


@pytest.mark.parametrize("head_type,dropped", [("image", ("text_", "audio_")), ("text", ("image_", "audio_")),
                                               ("al", ("image_",)), ("vl", ("audio_",))])
def test_state_dict_is_pruned_per_head_type(head_type, dropped):
    """one_peace_retrieval.py:133-150: keys of modalities the head does not use are dropped before the strict load."""
    sd = synth.make_state_dict(**TINY, seed=1)
    model = build(head_type, sd)
    names = [n for n, _ in model.named_parameters()]
    assert names and not any(any(d in n for d in dropped) for n in names)


def test_forward_without_cuda_fails_loudly():
    model = build("text", synth.make_state_dict(**TINY, modalities=("text",), seed=2))
    tok = torch.randint(4, 1000, (2, 8))
    with torch.no_grad(), pytest.raises(RuntimeError):
        model(src_tokens=tok, encoder_type="text")
