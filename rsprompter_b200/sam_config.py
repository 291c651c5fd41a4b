"""SAM architecture tables.

The reference resolves sizes with ``SamConfig.from_pretrained(hf_pretrain_name)`` (M:727,753,772,
890,908), which needs a ``config.json`` on disk or the network.  Offline we read a local
``config.json`` when ``hf_pretrain_name`` is a directory that has one, otherwise the arch is
parsed from the name exactly as MMPretrainSamVisionEncoder does (last ``-`` / ``_`` token,
M:824) and the published SAM sizes are used (VS:377-401).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field, replace


@dataclass(frozen=True)
class SamVisionArch:
    name: str
    hidden_size: int
    num_layers: int
    num_heads: int
    mlp_dim: int
    global_attn_indexes: tuple[int, ...]
    image_size: int = 1024
    patch_size: int = 16
    window_size: int = 14
    output_channels: int = 256
    layer_norm_eps: float = 1e-6
    num_pos_feats: int = 128
    scale: float | None = None  # SamPositionalEmbedding Gaussian scale (hidden_size // 2 in HF)
    output_hidden_states: bool = False

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    def pe_scale(self) -> float:
        return float(self.scale if self.scale is not None else self.hidden_size // 2)


VISION_ARCHS = {
    "base": SamVisionArch("base", 768, 12, 12, 3072, (2, 5, 8, 11)),
    "large": SamVisionArch("large", 1024, 24, 16, 4096, (5, 11, 17, 23)),
    "huge": SamVisionArch("huge", 1280, 32, 16, 5120, (7, 15, 23, 31)),
}


@dataclass(frozen=True)
class SamDecoderArch:
    hidden_size: int = 256
    num_heads: int = 8
    mlp_dim: int = 2048
    num_layers: int = 2
    attention_downsample_rate: int = 2
    num_multimask_outputs: int = 3
    iou_head_depth: int = 3
    iou_head_hidden_dim: int = 256
    layer_norm_eps: float = 1e-6
    mask_input_channels: int = 16  # prompt encoder SamMaskEmbedding


def parse_arch_name(hf_pretrain_name: str) -> str:
    """'facebook/sam-vit-huge' / 'work_dirs/sam_cache/sam_vit_base' -> 'huge' / 'base'."""
    tail = os.path.basename(os.path.normpath(hf_pretrain_name))
    tok = tail.split("-")[-1].split("_")[-1].lower()
    for k in VISION_ARCHS:
        if tok.startswith(k[0]) and (tok == k or tok == k[0]):
            return k
    for k in VISION_ARCHS:
        if k in tail.lower():
            return k
    raise ValueError(f"cannot infer SAM arch (base/large/huge) from {hf_pretrain_name!r}")


def vision_arch(hf_pretrain_name: str, extra_config: dict | None = None,
                img_size: int | None = None) -> SamVisionArch:
    arch = None
    cfg_path = os.path.join(hf_pretrain_name, "config.json")
    if os.path.isfile(cfg_path):
        with open(cfg_path, encoding="utf-8") as f:
            vc = json.load(f).get("vision_config", {})
        if vc:
            arch = SamVisionArch(
                name=parse_arch_name(hf_pretrain_name) if any(
                    k in hf_pretrain_name.lower() for k in VISION_ARCHS) else "custom",
                hidden_size=vc.get("hidden_size", 768), num_layers=vc.get("num_hidden_layers", 12),
                num_heads=vc.get("num_attention_heads", 12), mlp_dim=vc.get("mlp_dim", 3072),
                global_attn_indexes=tuple(vc.get("global_attn_indexes", (2, 5, 8, 11))),
                image_size=vc.get("image_size", 1024), patch_size=vc.get("patch_size", 16),
                window_size=vc.get("window_size", 14), output_channels=vc.get("output_channels", 256),
                layer_norm_eps=vc.get("layer_norm_eps", 1e-6), num_pos_feats=vc.get("num_pos_feats", 128),
                scale=vc.get("scale"))
    if arch is None:
        arch = VISION_ARCHS[parse_arch_name(hf_pretrain_name)]
    if extra_config:
        known = {k: v for k, v in extra_config.items() if k in SamVisionArch.__dataclass_fields__}
        arch = replace(arch, **known)
    if img_size is not None:
        arch = replace(arch, image_size=int(img_size))
    return arch


def decoder_arch(hf_pretrain_name: str | None = None, extra_config: dict | None = None) -> SamDecoderArch:
    arch = SamDecoderArch()
    if extra_config:
        known = {k: v for k, v in extra_config.items() if k in SamDecoderArch.__dataclass_fields__}
        arch = replace(arch, **known)
    return arch
