"""TEST INFRASTRUCTURE — not product code.

Loads the reference's OWN arithmetic files (``/root/reference/one_peace/models/**``,
``criterions/image_text_retrieval_loss.py``, ``optim/adam.py``) unmodified, under a minimal stub of the
parts of ``fairseq`` / ``timm`` they import (neither is installed here and there is no network; the
reference also fails to import on Python >= 3.11 because ``unify_model_config.py`` uses mutable
dataclass defaults — SURVEY.md §8c).  Only usable in the build container where ``/root/reference``
exists; it is used by ``oracle/make_golden.py`` to emit the fixtures under ``tests/golden/`` that pin
``oracle/restated.py``.  Nothing on the GPU box may import this module.

What is stubbed (each item cites what it stands in for, under /root/reference/fairseq/fairseq/):
  utils.softmax / log_softmax            utils.py:516-527   (F.(log_)softmax(x, dim, dtype=float32))
  utils.new_arange                       utils.py:707-714
  utils.get_available_activation_fns     utils.py (list of names; only used for a ChoiceEnum)
  modules.FairseqDropout                 modules/fairseq_dropout.py:16-27
  modules.LayerDropModuleList            modules/layer_drop.py (p == 0 in every ONE-PEACE config)
  modules.checkpoint_activations.checkpoint_wrapper, distributed.fsdp_wrap   -> identity
  models.BaseFairseqModel / FairseqEncoder / register_model                   -> nn.Module shells
  dataclass.FairseqDataclass / ChoiceEnum, models.transformer.EncDecBaseConfig (transformer_config.py:26-50)
  criterions.FairseqCriterion / register_criterion, metrics, optim.FairseqOptimizer / register_optimizer
  timm.models.layers.trunc_normal_       -> torch.nn.init.trunc_normal_
"""
import importlib
import importlib.util
import os
import re
import sys
import types
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = "/root/reference"
REF_PKG = os.path.join(REF_ROOT, "one_peace")


def reference_available():
    return os.path.isdir(REF_PKG)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _FairseqDropout(nn.Module):
    def __init__(self, p, module_name=None):
        super().__init__()
        self.p = p
        self.module_name = module_name
        self.apply_during_inference = False

    def forward(self, x, inplace: bool = False):
        if self.p > 0 and (self.training or self.apply_during_inference):
            return F.dropout(x, p=self.p, training=True, inplace=inplace)
        return x


class _LayerDropModuleList(nn.ModuleList):
    def __init__(self, p, modules=None):
        super().__init__(modules)
        self.p = p


class _BaseFairseqModel(nn.Module):
    def __init__(self):
        super().__init__()

    def set_num_updates(self, num_updates):
        pass

    def upgrade_state_dict_named(self, state_dict, name):
        pass


class _FairseqEncoder(nn.Module):
    def __init__(self, dictionary):
        super().__init__()
        self.dictionary = dictionary


@dataclass
class _FairseqDataclass:
    _name: Optional[str] = None


def _ChoiceEnum(choices):
    return str


@dataclass
class _EncDecBaseConfig(_FairseqDataclass):
    embed_path: Optional[str] = None
    embed_dim: Optional[int] = 512
    ffn_embed_dim: int = 2048
    layers: int = 6
    attention_heads: int = 8
    normalize_before: bool = False
    learned_pos: bool = False
    layerdrop: float = 0
    layers_to_keep: Optional[List[int]] = None


def _register(*a, **k):
    def deco(cls):
        return cls
    return deco


class _FairseqCriterion(nn.Module):
    def __init__(self, task):
        super().__init__()
        self.task = task


class _Dictionary:
    """Only len() and pad() are used on the path (adapter/text.py:41-43)."""

    def __init__(self, n=50264, pad=1):
        self._n, self._pad = n, pad

    def __len__(self):
        return self._n

    def pad(self):
        return self._pad


_installed = False


def install():
    """Install the stub modules and package shells; idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("/root/reference is not present: the reference stub only works in the build container")
    utils = _mod(
        "fairseq.utils",
        softmax=lambda x, dim, onnx_trace=False: F.softmax(x, dim=dim, dtype=torch.float32),
        log_softmax=lambda x, dim, onnx_trace=False: F.log_softmax(x, dim=dim, dtype=torch.float32),
        new_arange=lambda x, *size: torch.arange((size or x.size())[-1], device=x.device).expand(*(size or x.size())).contiguous(),
        get_available_activation_fns=lambda: ["relu", "gelu", "gelu_fast", "gelu_accurate", "tanh", "linear"],
    )
    metrics = _mod("fairseq.metrics", log_scalar=lambda *a, **k: None, log_derived=lambda *a, **k: None)
    fairseq = _mod("fairseq", utils=utils, metrics=metrics)
    fairseq.__path__ = []
    models = _mod("fairseq.models", BaseFairseqModel=_BaseFairseqModel, FairseqEncoder=_FairseqEncoder,
                  register_model=_register)
    models.__path__ = []
    _mod("fairseq.models.transformer", EncDecBaseConfig=_EncDecBaseConfig)
    modules = _mod("fairseq.modules", FairseqDropout=_FairseqDropout, LayerDropModuleList=_LayerDropModuleList)
    modules.__path__ = []
    _mod("fairseq.modules.fairseq_dropout", FairseqDropout=_FairseqDropout)
    _mod("fairseq.modules.checkpoint_activations", checkpoint_wrapper=lambda m, *a, **k: m)
    _mod("fairseq.distributed", fsdp_wrap=lambda m, *a, **k: m)
    _mod("fairseq.dataclass", FairseqDataclass=_FairseqDataclass, ChoiceEnum=_ChoiceEnum)
    _mod("fairseq.criterions", FairseqCriterion=_FairseqCriterion, register_criterion=_register)
    optim = _mod("fairseq.optim", FairseqOptimizer=object, register_optimizer=_register)
    optim.__path__ = []
    _mod("fairseq.optim.fused_adam", get_fused_adam_class=lambda: None)

    timm = _mod("timm")
    timm.__path__ = []
    tm = _mod("timm.models")
    tm.__path__ = []
    _mod("timm.models.layers", trunc_normal_=lambda t, mean=0.0, std=1.0, a=-2.0, b=2.0: nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b))

    # package shells so that the reference's __init__ side imports (hub_interface -> librosa, ...) are skipped
    for name, rel in [("one_peace", ""), ("one_peace.models", "models"), ("one_peace.models.one_peace", "models/one_peace"),
                      ("one_peace.criterions", "criterions"), ("one_peace.optim", "optim")]:
        m = _mod(name)
        m.__path__ = [os.path.join(REF_PKG, rel)]
    # Python >= 3.11: dataclass defaults `X: T = T()` -> field(default_factory=T) (the only patch applied)
    src = open(os.path.join(REF_PKG, "models", "unify_model_config.py")).read()
    src = re.sub(r"^(\s+\w+): (\w+Config) = (\w+Config)\(\)\s*$", r"\1: \2 = field(default_factory=\3)", src, flags=re.M)
    cfgmod = _mod("one_peace.models.unify_model_config")
    cfgmod.__file__ = os.path.join(REF_PKG, "models", "unify_model_config.py")
    exec(compile(src, cfgmod.__file__, "exec"), cfgmod.__dict__)
    _installed = True


def ref_module(name):
    """Import a reference module by dotted name, e.g. 'one_peace.models.one_peace.one_peace_retrieval'."""
    install()
    m = importlib.import_module(name)
    # force the pure-torch branches
    if name.endswith("components"):
        m.has_flash = False
    return m


def build_reference_retrieval(embed_dim=256, ffn=1024, layers=2, heads=4, head_type="val", patch_image_size=224,
                              text_bucket=256, image_bucket=16, audio_bucket=512, layer_scale_init=1e-6, seed=0,
                              vocab=50264):
    """Instantiate the reference OnePeaceRetrievalModel with the 4B config flags at a chosen width/depth."""
    install()
    comps = ref_module("one_peace.models.components")
    comps.has_flash = False
    mha = ref_module("one_peace.models.transformer.multihead_attention")
    mha.has_xformers = False
    retr = ref_module("one_peace.models.one_peace.one_peace_retrieval")
    cfg = retr.OnePeaceRetrievalConfig()
    enc = cfg.encoder
    enc.embed_dim, enc.ffn_embed_dim, enc.layers, enc.attention_heads = embed_dim, ffn, layers, heads
    enc.normalize_before, enc.learned_pos = True, True
    enc.drop_path_rate = 0.0
    enc.dropout = enc.attention_dropout = enc.activation_dropout = 0.0
    enc.magneto_scale_attn, enc.scale_attn, enc.scale_fc, enc.scale_heads = True, False, True, False
    enc.use_layer_scale, enc.layer_scale_init_value = True, layer_scale_init
    enc.text_adapter.bucket_size, enc.text_adapter.use_attn_bias = text_bucket, True
    enc.image_adapter.bucket_size, enc.image_adapter.use_attn_bias = image_bucket, True
    enc.image_adapter.rel_bucket_size = patch_image_size // 16
    enc.image_adapter.vision_encoder_type = "hmlp"
    enc.audio_adapter.bucket_size, enc.audio_adapter.use_attn_bias = audio_bucket, True
    torch.manual_seed(seed)
    model = retr.OnePeaceRetrievalModel(cfg, _Dictionary(vocab), head_type)
    model.eval()
    return model


def build_reference_pretrain(embed_dim=256, ffn=1024, layers=2, heads=4, dec_dim=128, dec_ffn=256, dec_layers=2, dec_heads=2,
                             res=64, vocab=1000, text_bucket=256, seed=0):
    """Instantiate the reference OnePeacePretrainModel (text + image experts) with the flags of pretrain_vl_3B.yaml:92-168 at a
    chosen width / depth (drop-path 0 so that the forward is deterministic)."""
    install()
    ref_module("one_peace.models.components").has_flash = False
    ref_module("one_peace.models.transformer.multihead_attention").has_xformers = False
    pre = ref_module("one_peace.models.one_peace.one_peace_pretrain")
    cfg = pre.OnePeacePretrainConfig()
    w = res // 16
    for part, (d, f, L, h) in (("encoder", (embed_dim, ffn, layers, heads)), ("decoder", (dec_dim, dec_ffn, dec_layers, dec_heads))):
        c = getattr(cfg, part)
        c.embed_dim, c.ffn_embed_dim, c.layers, c.attention_heads = d, f, L, h
        c.normalize_before, c.learned_pos = True, True
        c.drop_path_rate = 0.0
        c.dropout = c.attention_dropout = c.activation_dropout = 0.0
        c.magneto_scale_attn, c.scale_attn, c.scale_fc, c.scale_heads = True, False, True, False
        c.use_text_moe, c.use_image_moe, c.use_audio_moe = True, True, False
        c.checkpoint_activations = False
        enc = part == "encoder"
        c.use_layer_scale, c.layer_scale_init_value = enc, 1e-6
        c.text_adapter.bucket_size, c.text_adapter.use_attn_bias = text_bucket, enc
        c.image_adapter.bucket_size, c.image_adapter.rel_bucket_size, c.image_adapter.use_attn_bias = w, w, enc
        c.image_adapter.vision_encoder_type = "hmlp" if enc else "none"
    torch.manual_seed(seed)
    model = pre.OnePeacePretrainModel(cfg, _Dictionary(vocab))
    model.eval()
    return model


def build_reference_audio_pretrain(embed_dim=256, ffn=1024, layers=2, heads=4, dec_dim=128, dec_ffn=256, dec_layers=2, dec_heads=2,
                                   vocab=1000, text_bucket=256, audio_bucket=512, seed=0):
    """Instantiate the reference OnePeacePretrainModel (text + audio experts) with the flags of pretrain_al_3B.yaml:90-170 at a
    chosen width / depth: decoder audio adapter without feature extractor, abs_pos_type 'fixed', no attention bias."""
    install()
    ref_module("one_peace.models.components").has_flash = False
    ref_module("one_peace.models.transformer.multihead_attention").has_xformers = False
    pre = ref_module("one_peace.models.one_peace.one_peace_pretrain")
    cfg = pre.OnePeacePretrainConfig()
    for part, (d, f, L, h) in (("encoder", (embed_dim, ffn, layers, heads)), ("decoder", (dec_dim, dec_ffn, dec_layers, dec_heads))):
        c = getattr(cfg, part)
        c.embed_dim, c.ffn_embed_dim, c.layers, c.attention_heads = d, f, L, h
        c.normalize_before, c.learned_pos = True, True
        c.drop_path_rate = 0.0
        c.dropout = c.attention_dropout = c.activation_dropout = 0.0
        c.magneto_scale_attn, c.scale_attn, c.scale_fc, c.scale_heads = True, False, True, False
        c.use_text_moe, c.use_image_moe, c.use_audio_moe = True, False, True
        c.checkpoint_activations = False
        enc = part == "encoder"
        c.use_layer_scale, c.layer_scale_init_value = enc, 1e-6
        c.text_adapter.bucket_size, c.text_adapter.use_attn_bias = text_bucket, enc
        c.audio_adapter.bucket_size, c.audio_adapter.use_attn_bias = audio_bucket, enc
        if not enc:
            c.audio_adapter.feature_encoder_spec = None
            c.audio_adapter.abs_pos_type = "fixed"
    torch.manual_seed(seed)
    model = pre.OnePeacePretrainModel(cfg, _Dictionary(vocab))
    model.eval()
    return model
