"""Seeded random checkpoints with the reference's parameter names and shapes.

No SAM / RSPrompter weights exist offline, so benchmarks and parity tests run on random
weights of the exact architecture (HF ``pytorch_model.bin`` key names
with the ``vision_encoder.`` prefix stripped as M:783 does).  Relative-position tables and the
absolute position embedding are drawn non-zero (HF zero-initialises them, HF:1008-1014, which
would hide every rel-pos bug).  The same dicts are loaded into the oracle (HF modules /
restatement) and into the B200 modules, so parity compares arithmetic, not initialisation.
"""
from __future__ import annotations

import math

import torch

from .sam_config import SamDecoderArch, SamVisionArch


def _randn(gen: torch.Generator, *shape: int, std: float = 0.02) -> torch.Tensor:
    return torch.randn(*shape, generator=gen, dtype=torch.float32) * std


def _linear(sd: dict, gen: torch.Generator, prefix: str, out_f: int, in_f: int, bias: bool = True,
            std: float | None = None) -> None:
    std = std if std is not None else 1.0 / math.sqrt(in_f)
    sd[prefix + ".weight"] = _randn(gen, out_f, in_f, std=std)
    if bias:
        sd[prefix + ".bias"] = _randn(gen, out_f, std=0.02)


def _norm(sd: dict, gen: torch.Generator, prefix: str, c: int) -> None:
    sd[prefix + ".weight"] = 1.0 + _randn(gen, c, std=0.05)
    sd[prefix + ".bias"] = _randn(gen, c, std=0.05)


def vision_encoder_state_dict(arch: SamVisionArch, seed: int = 0) -> dict[str, torch.Tensor]:
    """Keys of HF ``SamVisionEncoder`` (= RSSamVisionEncoder.vision_encoder.*)."""
    gen = torch.Generator().manual_seed(seed)
    D, hd, g = arch.hidden_size, arch.head_dim, arch.grid
    sd: dict[str, torch.Tensor] = {}
    sd["pos_embed"] = _randn(gen, 1, g, g, D, std=0.02)
    sd["patch_embed.projection.weight"] = _randn(gen, D, 3, arch.patch_size, arch.patch_size,
                                                 std=1.0 / math.sqrt(3 * arch.patch_size ** 2))
    sd["patch_embed.projection.bias"] = _randn(gen, D, std=0.02)
    for i in range(arch.num_layers):
        p = f"layers.{i}."
        S = g if i in arch.global_attn_indexes else arch.window_size
        _norm(sd, gen, p + "layer_norm1", D)
        sd[p + "attn.rel_pos_h"] = _randn(gen, 2 * S - 1, hd, std=0.1)
        sd[p + "attn.rel_pos_w"] = _randn(gen, 2 * S - 1, hd, std=0.1)
        _linear(sd, gen, p + "attn.qkv", 3 * D, D)
        _linear(sd, gen, p + "attn.proj", D, D, std=0.5 / math.sqrt(D))
        _norm(sd, gen, p + "layer_norm2", D)
        _linear(sd, gen, p + "mlp.lin1", arch.mlp_dim, D)
        _linear(sd, gen, p + "mlp.lin2", D, arch.mlp_dim, std=0.5 / math.sqrt(arch.mlp_dim))
    C = arch.output_channels
    sd["neck.conv1.weight"] = _randn(gen, C, D, 1, 1, std=1.0 / math.sqrt(D))
    _norm(sd, gen, "neck.layer_norm1", C)
    sd["neck.conv2.weight"] = _randn(gen, C, C, 3, 3, std=1.0 / math.sqrt(9 * C))
    _norm(sd, gen, "neck.layer_norm2", C)
    return sd


def mask_decoder_state_dict(arch: SamDecoderArch | None = None, seed: int = 1) -> dict[str, torch.Tensor]:
    """Keys of HF ``SamMaskDecoder`` (= RSSamMaskDecoder.mask_decoder.*)."""
    arch = arch or SamDecoderArch()
    gen = torch.Generator().manual_seed(seed)
    C = arch.hidden_size
    sd: dict[str, torch.Tensor] = {}
    sd["iou_token.weight"] = _randn(gen, 1, C, std=0.5)
    sd["mask_tokens.weight"] = _randn(gen, arch.num_multimask_outputs + 1, C, std=0.5)

    def attn(prefix: str, internal: int) -> None:
        for n in ("q_proj", "k_proj", "v_proj"):
            _linear(sd, gen, f"{prefix}.{n}", internal, C)
        _linear(sd, gen, f"{prefix}.out_proj", C, internal)

    for i in range(arch.num_layers):
        p = f"transformer.layers.{i}."
        attn(p + "self_attn", C)
        attn(p + "cross_attn_token_to_image", C // arch.attention_downsample_rate)
        attn(p + "cross_attn_image_to_token", C // arch.attention_downsample_rate)
        _linear(sd, gen, p + "mlp.lin1", arch.mlp_dim, C)
        _linear(sd, gen, p + "mlp.lin2", C, arch.mlp_dim)
        for k in range(1, 5):
            _norm(sd, gen, p + f"layer_norm{k}", C)
    attn("transformer.final_attn_token_to_image", C // arch.attention_downsample_rate)
    _norm(sd, gen, "transformer.layer_norm_final_attn", C)
    sd["upscale_conv1.weight"] = _randn(gen, C, C // 4, 2, 2, std=1.0 / math.sqrt(C))
    sd["upscale_conv1.bias"] = _randn(gen, C // 4, std=0.02)
    _norm(sd, gen, "upscale_layer_norm", C // 4)
    sd["upscale_conv2.weight"] = _randn(gen, C // 4, C // 8, 2, 2, std=1.0 / math.sqrt(C // 4))
    sd["upscale_conv2.bias"] = _randn(gen, C // 8, std=0.02)
    for i in range(arch.num_multimask_outputs + 1):
        p = f"output_hypernetworks_mlps.{i}."
        _linear(sd, gen, p + "proj_in", C, C)
        _linear(sd, gen, p + "layers.0", C, C)
        _linear(sd, gen, p + "proj_out", C // 8, C)
    _linear(sd, gen, "iou_prediction_head.proj_in", arch.iou_head_hidden_dim, C)
    for k in range(arch.iou_head_depth - 2):
        _linear(sd, gen, f"iou_prediction_head.layers.{k}", arch.iou_head_hidden_dim, arch.iou_head_hidden_dim)
    _linear(sd, gen, "iou_prediction_head.proj_out", arch.num_multimask_outputs + 1, arch.iou_head_hidden_dim)
    return sd


def prompt_encoder_state_dict(arch: SamDecoderArch | None = None, seed: int = 2) -> dict[str, torch.Tensor]:
    """The members of HF ``SamPromptEncoder`` the path uses (M:305-307,1635): no_mask_embed, mask_embed."""
    arch = arch or SamDecoderArch()
    gen = torch.Generator().manual_seed(seed)
    C, mc = arch.hidden_size, arch.mask_input_channels
    sd: dict[str, torch.Tensor] = {}
    sd["no_mask_embed.weight"] = _randn(gen, 1, C, std=0.5)
    sd["mask_embed.conv1.weight"] = _randn(gen, mc // 4, 1, 2, 2, std=0.5)
    sd["mask_embed.conv1.bias"] = _randn(gen, mc // 4, std=0.1)
    _norm(sd, gen, "mask_embed.layer_norm1", mc // 4)
    sd["mask_embed.conv2.weight"] = _randn(gen, mc, mc // 4, 2, 2, std=0.25)
    sd["mask_embed.conv2.bias"] = _randn(gen, mc, std=0.1)
    _norm(sd, gen, "mask_embed.layer_norm2", mc)
    sd["mask_embed.conv3.weight"] = _randn(gen, C, mc, 1, 1, std=0.25)
    sd["mask_embed.conv3.bias"] = _randn(gen, C, std=0.1)
    return sd


def positional_embedding_state_dict(arch: SamVisionArch, seed: int = 3) -> dict[str, torch.Tensor]:
    """``shared_image_embedding.positional_embedding`` (2, num_pos_feats) = scale * randn (HF:549-550).

    A scale of 1.0 keeps the synthetic Fourier features smooth enough to be a meaningful
    numerical test (HF's default 384 for ViT-B turns sin/cos of 2*pi*x into noise)."""
    gen = torch.Generator().manual_seed(seed)
    return {"positional_embedding": _randn(gen, 2, arch.num_pos_feats, std=1.0)}
