"""Config tree with the reference's field names (models/unify_model_config.py:14-219), written with
``default_factory`` so it also imports on Python >= 3.11 (the reference file does not)."""
from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class TextAdapterConfig:
    bucket_size: int = 256
    layernorm_embedding: bool = False
    add_type_embedding: bool = False
    shrink_alpha: float = 1.0
    dropout: float = 0.0
    use_attn_bias: bool = False


@dataclass
class ImageAdapterConfig:
    bucket_size: int = 16
    rel_bucket_size: int = 16
    layernorm_embedding: bool = False
    add_type_embedding: bool = False
    vision_encoder_type: str = "hmlp"
    shrink_alpha: float = 1.0
    dropout: float = 0.0
    use_attn_bias: bool = False


@dataclass
class AudioAdapterConfig:
    feature_embed_dim: int = 512
    feature_encoder_spec: Optional[str] = "[(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512,2,2)] + [(512,2,2)]"
    abs_pos_type: str = "conv"
    conv_pos_depth: int = 5
    conv_pos_width: int = 95
    conv_pos_groups: int = 16
    conv_pos_pre_ln: bool = False
    bucket_size: int = 256
    layernorm_embedding: bool = False
    add_type_embedding: bool = False
    shrink_alpha: float = 1.0
    dropout: float = 0.0
    use_attn_bias: bool = False
    conv_bias: bool = False
    freeze_extractor: bool = False


@dataclass
class AdjustEncDecConfig:
    # fairseq EncDecBaseConfig fields (fairseq/models/transformer/transformer_config.py:26-50)
    embed_path: Optional[str] = None
    embed_dim: Optional[int] = 512
    ffn_embed_dim: int = 2048
    layers: int = 6
    attention_heads: int = 8
    normalize_before: bool = False
    learned_pos: bool = False
    layerdrop: float = 0
    layers_to_keep: Optional[List[int]] = None
    # ONE-PEACE additions
    text_adapter: TextAdapterConfig = field(default_factory=TextAdapterConfig)
    image_adapter: ImageAdapterConfig = field(default_factory=ImageAdapterConfig)
    audio_adapter: AudioAdapterConfig = field(default_factory=AudioAdapterConfig)
    drop_path_rate: float = 0.0
    magneto_scale_attn: bool = False
    scale_attn: bool = True
    scale_fc: bool = True
    scale_heads: bool = True
    use_text_moe: bool = True
    use_image_moe: bool = True
    use_audio_moe: bool = True
    use_layer_scale: bool = True
    layer_scale_init_value: float = 1e-2
    activation_fn: str = "gelu"
    dropout: float = 0.1
    attention_dropout: float = 0.0
    activation_dropout: float = 0.0
    max_positions: int = 1024
    checkpoint_activations: bool = False
    fsdp_checkpoint_wrap_layer_preserve_frequency: Optional[int] = 1
    fsdp_checkpoint_wrap_layer_skip_frequency: Optional[int] = 1000
    offload_activations: bool = False


@dataclass
class UnifyModelConfig:
    _name: Optional[str] = None
    encoder: AdjustEncDecConfig = field(default_factory=AdjustEncDecConfig)
    decoder: AdjustEncDecConfig = field(default_factory=AdjustEncDecConfig)


def one_peace_4b_decoder_config(embed_dim=768, ffn_embed_dim=2048, layers=2, attention_heads=12, patch_image_size=256):
    """Decoder section of run_scripts/pretrain/pretrain_vl_3B.yaml:127-168 (pretrain_al_3B.yaml:129-170 for the audio adapter:
    no feature extractor, learned 'fixed' positions): no LayerScale, no relative-position bias, no stems."""
    c = AdjustEncDecConfig(embed_dim=embed_dim, ffn_embed_dim=ffn_embed_dim, layers=layers, attention_heads=attention_heads,
                           normalize_before=True, learned_pos=True, drop_path_rate=0.0, dropout=0.0, attention_dropout=0.0,
                           magneto_scale_attn=True, scale_attn=False, scale_fc=True, scale_heads=False, use_layer_scale=False,
                           layer_scale_init_value=1e-6, use_audio_moe=False)
    c.text_adapter = TextAdapterConfig(bucket_size=256, use_attn_bias=False)
    c.image_adapter = ImageAdapterConfig(bucket_size=patch_image_size // 16, rel_bucket_size=patch_image_size // 16,
                                         vision_encoder_type="none", use_attn_bias=False)
    c.audio_adapter = AudioAdapterConfig(feature_encoder_spec=None, abs_pos_type="fixed", use_attn_bias=False)
    return c


def one_peace_4b_encoder_config(layers=40, embed_dim=1536, ffn_embed_dim=6144, attention_heads=24,
                                patch_image_size=224):
    """Encoder section of run_scripts/finetune_3B.yaml:76-132 (the "4B" config)."""
    c = AdjustEncDecConfig(embed_dim=embed_dim, ffn_embed_dim=ffn_embed_dim, layers=layers,
                           attention_heads=attention_heads, normalize_before=True, learned_pos=True,
                           drop_path_rate=0.0, dropout=0.0, attention_dropout=0.0, magneto_scale_attn=True,
                           scale_attn=False, scale_fc=True, scale_heads=False, use_layer_scale=True,
                           layer_scale_init_value=1e-6)
    c.text_adapter = TextAdapterConfig(bucket_size=256, use_attn_bias=True)
    c.image_adapter = ImageAdapterConfig(bucket_size=16, rel_bucket_size=patch_image_size // 16, use_attn_bias=True)
    c.audio_adapter = AudioAdapterConfig(bucket_size=512, use_attn_bias=True)
    return c
