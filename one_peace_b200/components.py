"""Parameter containers with the reference's factory names (models/components.py:23-44).

They are plain torch modules used only to HOLD parameters under the reference's names (checkpoint
compatibility, §8a); the arithmetic runs in the sm_100a kernels.
"""
import torch
import torch.nn as nn


def trunc_normal_(tensor, mean=0.0, std=0.02):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=-std, b=std)


def LayerNorm(normalized_shape, eps=1e-5, elementwise_affine=True):
    return torch.nn.LayerNorm(normalized_shape, eps, elementwise_affine)


def Linear(in_features, out_features, bias=True):
    m = nn.Linear(in_features, out_features, bias)
    nn.init.xavier_uniform_(m.weight)
    if bias:
        nn.init.constant_(m.bias, 0.0)
    return m


def Embedding(num_embeddings, embedding_dim, padding_idx=None, zero_init=False):
    m = nn.Embedding(num_embeddings, embedding_dim, padding_idx=padding_idx)
    nn.init.normal_(m.weight, mean=0, std=embedding_dim ** -0.5)
    if padding_idx is not None:
        nn.init.constant_(m.weight[padding_idx], 0)
    if zero_init:
        nn.init.constant_(m.weight, 0)
    return m


def f32(t):
    """Contiguous fp32 copy of a parameter for kernel consumption."""
    return t.detach().to(torch.float32).contiguous()


def bf16(t):
    return t.detach().to(torch.bfloat16).contiguous()


class PackCache:
    """Caches kernel-ready (bf16 / fp32, re-laid-out) copies of a module's parameters and rebuilds them
    when any source parameter was modified in place (optimizer step, load_state_dict) or moved."""

    def __init__(self):
        self._key = None
        self._pack = None

    def get(self, params, builder):
        key = tuple((p.data_ptr(), p._version, p.dtype, p.device) for p in params)
        if key != self._key:
            self._pack = builder()
            self._key = key
        return self._pack
