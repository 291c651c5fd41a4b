"""The third-party modules the reference calls, instantiated on seeded weights (CPU, fp32).

Reference call sites: M:14-16 (imports), M:730,756,775,893,911 (construction), M:94,98,369,1685
(calls).  transformers 5.5.0 differs from the pinned 4.38.1 in ways listed in SURVEY.md 8(c);
the shims here are: eager attention is forced, the decoder's 2-tuple is used as is, and
``SamMaskEmbedding`` is built directly from the prompt-encoder config.
"""
from __future__ import annotations

import torch


def _vision_config(arch, output_hidden_states: bool = True):
    from transformers import SamVisionConfig
    cfg = SamVisionConfig(
        hidden_size=arch.hidden_size, output_channels=arch.output_channels,
        num_hidden_layers=arch.num_layers, num_attention_heads=arch.num_heads,
        image_size=arch.image_size, patch_size=arch.patch_size, layer_norm_eps=arch.layer_norm_eps,
        window_size=arch.window_size, global_attn_indexes=list(arch.global_attn_indexes),
        num_pos_feats=arch.num_pos_feats, mlp_dim=arch.mlp_dim)
    cfg._attn_implementation = "eager"
    cfg.output_hidden_states = output_hidden_states
    return cfg


def build_vision_encoder(arch, state_dict):
    from transformers.models.sam.modeling_sam import SamVisionEncoder
    m = SamVisionEncoder(_vision_config(arch))
    missing, unexpected = m.load_state_dict(state_dict, strict=True), None
    m.eval()
    return m


def run_vision_encoder(model, pixel_values):
    """-> (embeddings [B,C,g,g], tuple of L+1 hidden states [B,g,g,D]) as the detectors unpack (M:99-101)."""
    with torch.no_grad():
        out = model(pixel_values, output_hidden_states=True)
    return out[0], out[1]


def _decoder_config(arch):
    from transformers import SamMaskDecoderConfig
    cfg = SamMaskDecoderConfig(
        hidden_size=arch.hidden_size, mlp_dim=arch.mlp_dim, num_hidden_layers=arch.num_layers,
        num_attention_heads=arch.num_heads, attention_downsample_rate=arch.attention_downsample_rate,
        num_multimask_outputs=arch.num_multimask_outputs, iou_head_depth=arch.iou_head_depth,
        iou_head_hidden_dim=arch.iou_head_hidden_dim, layer_norm_eps=arch.layer_norm_eps)
    cfg._attn_implementation = "eager"
    return cfg


def build_mask_decoder(arch, state_dict):
    from transformers.models.sam.modeling_sam import SamMaskDecoder
    m = SamMaskDecoder(_decoder_config(arch))
    m.load_state_dict(state_dict, strict=True)
    m.eval()
    return m


def build_mask_embedding(arch, state_dict):
    """HF SamMaskEmbedding (prompt_encoder.mask_embed, M:305) from 'mask_embed.*' keys."""
    from transformers import SamPromptEncoderConfig
    from transformers.models.sam.modeling_sam import SamMaskEmbedding
    cfg = SamPromptEncoderConfig(hidden_size=arch.hidden_size, mask_input_channels=arch.mask_input_channels,
                                 layer_norm_eps=arch.layer_norm_eps)
    m = SamMaskEmbedding(cfg)
    m.load_state_dict({k[len("mask_embed."):]: v for k, v in state_dict.items()
                       if k.startswith("mask_embed.")}, strict=True)
    m.eval()
    return m


def build_positional_embedding(vision_arch, state_dict):
    from transformers.models.sam.modeling_sam import SamPositionalEmbedding
    m = SamPositionalEmbedding(_vision_config(vision_arch))
    m.load_state_dict(state_dict, strict=True)
    m.eval()
    return m
