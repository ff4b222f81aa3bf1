"""Drop-in for ``AudioAdapter`` (models/adapter/audio.py:35-311) — placeholder until the conv stack lands."""
import torch


class AudioAdapter(torch.nn.Module):
    def __init__(self, cfg, embed_dim, attention_heads, num_layers=None):
        super().__init__()
        raise NotImplementedError("audio adapter: under construction")
