"""Pins the oracle restatement (oracle/restate.py) against the third-party modules the reference
actually calls (HF transformers SAM classes, eager attention), on seeded weights, CPU fp32."""
import dataclasses

import pytest
import torch

from oracle import hf_ref, restate
from rsprompter_b200 import synthetic
from rsprompter_b200.sam_config import SamDecoderArch, SamVisionArch

TINY = SamVisionArch("tiny", hidden_size=128, num_layers=3, num_heads=2, mlp_dim=256,
                     global_attn_indexes=(1,), image_size=256)


def test_vit_encoder_restatement_matches_hf():
    torch.manual_seed(0)
    sd = synthetic.vision_encoder_state_dict(TINY, seed=5)
    x = torch.randn(2, 3, 256, 256)
    hf = hf_ref.build_vision_encoder(TINY, sd)
    emb_hf, hid_hf = hf_ref.run_vision_encoder(hf, x)
    emb, hid = restate.vit_encoder(sd, TINY, x)
    assert len(hid) == len(hid_hf) == TINY.num_layers + 1
    for a, b in zip(hid, hid_hf):
        assert a.shape == b.shape
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(emb, emb_hf, rtol=1e-4, atol=1e-4)


def test_rel_pos_is_not_degenerate():
    sd = synthetic.vision_encoder_state_dict(TINY, seed=5)
    sd2 = dict(sd)
    sd2["layers.0.attn.rel_pos_h"] = torch.zeros_like(sd["layers.0.attn.rel_pos_h"])
    x = torch.randn(1, 3, 256, 256)
    a, _ = restate.vit_encoder(sd, TINY, x)
    b, _ = restate.vit_encoder(sd2, TINY, x)
    assert (a - b).abs().max() > 1e-4


@pytest.mark.parametrize("multimask", [False, True])
def test_mask_decoder_restatement_matches_hf(multimask):
    torch.manual_seed(1)
    arch = SamDecoderArch()
    sd = synthetic.mask_decoder_state_dict(arch, seed=7)
    dec = hf_ref.build_mask_decoder(arch, sd)
    N, h = 3, 16
    emb = torch.randn(N, 256, h, h)
    pe = torch.randn(N, 256, h, h)
    sparse = torch.randn(N, 1, 5, 256)
    dense = torch.randn(N, 256, h, h)
    with torch.no_grad():
        m_hf, iou_hf = dec(image_embeddings=emb, image_positional_embeddings=pe,
                           sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                           multimask_output=multimask)
    m, iou = restate.mask_decoder(sd, arch, emb, pe, sparse, dense, multimask)
    assert m.shape == m_hf.shape and iou.shape == iou_hf.shape
    torch.testing.assert_close(m, m_hf, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(iou, iou_hf, rtol=1e-4, atol=1e-4)


def test_positional_embedding_matches_hf():
    sd = synthetic.positional_embedding_state_dict(TINY, seed=3)
    mod = hf_ref.build_positional_embedding(TINY, sd)
    size = 16
    grid = torch.ones(size, size)
    y = (grid.cumsum(0) - 0.5) / size
    x = (grid.cumsum(1) - 0.5) / size
    with torch.no_grad():
        ref = mod(torch.stack([x, y], dim=-1)).permute(2, 0, 1)[None]
    got = restate.image_wide_positional_embedding(sd["positional_embedding"], size)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


def test_mask_embedding_matches_hf():
    arch = SamDecoderArch()
    sd = synthetic.prompt_encoder_state_dict(arch, seed=2)
    mod = hf_ref.build_mask_embedding(arch, sd)
    x = torch.randn(4, 1, 64, 64)
    with torch.no_grad():
        ref = mod(x)
    got = restate.sam_mask_embedding(sd, x)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


# ---- query-path bricks whose mmcv originals are absent: pinned to independent implementations of the same operators
@pytest.mark.parametrize("E", [128, 256])
def test_ms_deform_attn_restatement_matches_hf_module(E):
    """oracle.restate_query.ms_deform_attn (mmcv MultiScaleDeformableAttention, batch_first, identity residual) against
    transformers' Mask2FormerPixelDecoderEncoderMultiscaleDeformableAttention - a separate port of the Deformable-DETR
    operator with the same parameter names - on seeded weights."""
    from transformers.models.mask2former.modeling_mask2former import (
        Mask2FormerPixelDecoderEncoderMultiscaleDeformableAttention as HFAttn)
    from oracle import restate_query as rq
    torch.manual_seed(E)
    heads, L, P, B = 8, 3, 4, 2
    shapes = [(6, 5), (12, 10), (24, 20)]
    nq = sum(h * w for h, w in shapes)
    mod = HFAttn(E, heads, L, P).eval()
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.05)
        mod.sampling_offsets.bias.normal_(0, 1.5)
        mod.attention_weights.weight.normal_(0, 0.1)
        mod.attention_weights.bias.normal_(0, 0.5)
    sd = {"a." + k: v.detach() for k, v in mod.state_dict().items()}
    x, pos = torch.randn(B, nq, E), torch.randn(B, nq, E)
    refs = []
    for h, w in shapes:
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        refs.append(torch.stack([(xs.reshape(-1) + 0.5) / w, (ys.reshape(-1) + 0.5) / h], dim=-1))
    ref = torch.cat(refs)[None, :, None].repeat(B, 1, L, 1)
    with torch.no_grad():
        want, _ = mod(x, encoder_hidden_states=x, position_embeddings=pos, reference_points=ref, spatial_shapes_list=shapes)
        got = rq.ms_deform_attn(sd, "a.", x, pos, ref, shapes, heads, P)
    torch.testing.assert_close(got - x, want, rtol=1e-4, atol=1e-4)          # the restatement adds the identity


@pytest.mark.parametrize("E,masked", [(128, True), (256, True), (256, False)])
def test_mha_restatement_matches_torch_multihead_attention(E, masked):
    """oracle.restate_query._mha (mmcv MultiheadAttention wrapper: q + query_pos, k + key_pos, identity + out) against
    torch.nn.MultiheadAttention(batch_first=True) with a boolean attn_mask [B*heads, nq, nk]."""
    from oracle import restate_query as rq
    torch.manual_seed(E + masked)
    heads, B, nq, nk = 8, 2, 7, 33
    mha = torch.nn.MultiheadAttention(E, heads, batch_first=True).eval()
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.1)
        mha.out_proj.bias.normal_(0, 0.1)
    sd = {"m.attn." + k: v.detach() for k, v in mha.state_dict().items()}
    q, qp = torch.randn(B, nq, E), torch.randn(B, nq, E)
    k, kp = torch.randn(B, nk, E), torch.randn(B, nk, E)
    am = None
    if masked:
        am = torch.rand(B, 1, nq, nk) < 0.5
        am[:, :, :, 0] = False                                            # no fully masked row
        am = am.repeat(1, heads, 1, 1).flatten(0, 1)
    with torch.no_grad():
        want = q + mha(q + qp, k + kp, k, attn_mask=am, need_weights=False)[0]
        got = rq._mha(sd, "m.", q, k, k, qp, kp, am, heads)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("E,F", [(128, 512), (256, 1024)])
def test_pixel_decoder_restatement_matches_hf_module(E, F):
    """oracle.restate_query.pixel_decoder (mmdet MSDeformAttnPixelDecoder.forward, msdeformattn_pixel_decoder.py:144-246)
    against transformers' Mask2FormerPixelDecoder - an independent port of the same Mask2Former module - configured
    like the RSPrompter configs (5 input levels, strides 4..64, 3 deformable-encoder levels, 2 FPN levels, GroupNorm 32)
    and given the same seeded weights through the name map below: mask features and the three memories."""
    from transformers import Mask2FormerConfig
    from transformers.models.mask2former.modeling_mask2former import Mask2FormerPixelDecoder
    from oracle import restate_query as rq
    torch.manual_seed(E)
    C = 256
    cfg = Mask2FormerConfig(feature_size=E, mask_feature_size=C, encoder_layers=3, num_attention_heads=8,
                            encoder_feedforward_dim=F, feature_strides=[4, 8, 16, 32, 64], common_stride=4, dropout=0.0)
    mod = Mask2FormerPixelDecoder(cfg, feature_channels=[C] * 5).eval()
    with torch.no_grad():
        for prm in mod.parameters():
            prm.normal_(0, 0.05)
        for n, prm in mod.named_parameters():
            if n.endswith(("1.weight",)) and prm.dim() == 1:          # norm scales around 1
                prm.add_(1.0)
        mod.level_embed.normal_(0, 0.5)
        for l in mod.encoder.layers:
            l.self_attn.sampling_offsets.bias.normal_(0, 1.5)
            for ln in (l.self_attn_layer_norm, l.final_layer_norm):
                ln.weight.fill_(1.0).add_(0.1 * torch.randn(E))
    hf = mod.state_dict()
    sd = {}

    def put(dst, src):
        for suf in ("weight", "bias"):
            if f"{src}.{suf}" in hf:
                sd[f"{dst}.{suf}"] = hf[f"{src}.{suf}"]

    for i in range(3):                                               # lowest resolution first in both
        put(f"input_convs.{i}.conv", f"input_projections.{i}.0")
        put(f"input_convs.{i}.gn", f"input_projections.{i}.1")
    sd["level_encoding.weight"] = hf["level_embed"]
    for l in range(3):
        for nm in ("sampling_offsets", "attention_weights", "value_proj", "output_proj"):
            put(f"encoder.layers.{l}.self_attn.{nm}", f"encoder.layers.{l}.self_attn.{nm}")
        put(f"encoder.layers.{l}.norms.0", f"encoder.layers.{l}.self_attn_layer_norm")
        put(f"encoder.layers.{l}.ffn.layers.0.0", f"encoder.layers.{l}.fc1")
        put(f"encoder.layers.{l}.ffn.layers.1", f"encoder.layers.{l}.fc2")
        put(f"encoder.layers.{l}.norms.1", f"encoder.layers.{l}.final_layer_norm")
    for i in range(2):   # mmdet's forward indexes lateral_convs / output_convs by the input level (0 = stride 4)
        put(f"lateral_convs.{i}.conv", f"adapter_{i + 1}.0")
        put(f"lateral_convs.{i}.gn", f"adapter_{i + 1}.1")
        put(f"output_convs.{i}.conv", f"layer_{i + 1}.0")
        put(f"output_convs.{i}.gn", f"layer_{i + 1}.1")
    put("mask_feature", "mask_projection")
    feats = [torch.randn(2, C, 128 // s, 160 // s) for s in (4, 8, 16, 32, 64)]
    with torch.no_grad():
        want = mod(feats)
        mf, mems = rq.pixel_decoder(sd, feats, "")
    torch.testing.assert_close(mf, want.mask_features, rtol=1e-3, atol=1e-3)
    assert len(mems) == len(want.multi_scale_features) == 3
    for a, b in zip(mems, want.multi_scale_features):
        torch.testing.assert_close(a, b, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("n_layers", [3, 9])
def test_stock_mask2former_decoder_restatement_matches_hf_module(n_layers):
    """oracle.restate_query.stock_mask2former_decoder (mmdet Mask2FormerHead.forward behind the pixel decoder:
    mask2former_head.py:340-460, layers/transformer/mask2former_layers.py:113-135) against transformers'
    Mask2FormerTransformerModule (masked-attention decoder + mask predictor; an independent port of the same module)
    on the same seeded weights: the last layer's mask logits and its post-norm query features (-> class logits).
    Both run fp32, so the thresholded-mask feedback sees the same bits."""
    from transformers import Mask2FormerConfig
    from transformers.models.mask2former.modeling_mask2former import Mask2FormerTransformerModule
    from oracle import restate_query as rq
    torch.manual_seed(100 + n_layers)
    E, C, nq, Fd, ncls = 256, 256, 12, 512, 10
    cfg = Mask2FormerConfig(hidden_dim=E, feature_size=E, mask_feature_size=C, num_queries=nq, decoder_layers=n_layers + 1,
                            num_attention_heads=8, dim_feedforward=Fd, dropout=0.0, enforce_input_projection=False)
    mod = Mask2FormerTransformerModule(in_features=E, config=cfg).eval()
    with torch.no_grad():
        for n, prm in mod.named_parameters():
            if "layer_norm" in n or "layernorm" in n:
                prm.copy_(1.0 + 0.1 * torch.randn_like(prm) if n.endswith("weight") else 0.1 * torch.randn_like(prm))
            elif prm.dim() >= 2:
                prm.normal_(0, 1.3 / prm.shape[-1] ** 0.5)
            else:
                prm.normal_(0, 0.05)
        mod.queries_embedder.weight.normal_(0, 1.0)
        mod.queries_features.weight.normal_(0, 1.0)
        mod.level_embed.weight.normal_(0, 0.5)
    hf = mod.state_dict()
    sd = {"query_embed.weight": hf["queries_embedder.weight"], "query_feat.weight": hf["queries_features.weight"],
          "level_embed.weight": hf["level_embed.weight"],
          "transformer_decoder.post_norm.weight": hf["decoder.layernorm.weight"],
          "transformer_decoder.post_norm.bias": hf["decoder.layernorm.bias"]}
    for i in range(n_layers):
        h, m = f"decoder.layers.{i}.", f"transformer_decoder.layers.{i}."
        for suf in ("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias"):
            sd[f"{m}cross_attn.attn.{suf}"] = hf[f"{h}cross_attn.{suf}"]
        sd[m + "self_attn.attn.in_proj_weight"] = torch.cat([hf[h + f"self_attn.{x}_proj.weight"] for x in "qkv"])
        sd[m + "self_attn.attn.in_proj_bias"] = torch.cat([hf[h + f"self_attn.{x}_proj.bias"] for x in "qkv"])
        for suf in ("weight", "bias"):
            sd[f"{m}self_attn.attn.out_proj.{suf}"] = hf[f"{h}self_attn.out_proj.{suf}"]
            sd[f"{m}norms.0.{suf}"] = hf[f"{h}cross_attn_layer_norm.{suf}"]
            sd[f"{m}norms.1.{suf}"] = hf[f"{h}self_attn_layer_norm.{suf}"]
            sd[f"{m}norms.2.{suf}"] = hf[f"{h}final_layer_norm.{suf}"]
            sd[f"{m}ffn.layers.0.0.{suf}"] = hf[f"{h}fc1.{suf}"]
            sd[f"{m}ffn.layers.1.{suf}"] = hf[f"{h}fc2.{suf}"]
    for j, k in enumerate((0, 2, 4)):
        for suf in ("weight", "bias"):
            sd[f"mask_embed.{k}.{suf}"] = hf[f"decoder.mask_predictor.mask_embedder.{j}.0.{suf}"]
    sd["cls_embed.weight"] = torch.randn(ncls + 1, E) * 0.1
    sd["cls_embed.bias"] = torch.randn(ncls + 1) * 0.1
    B = 2
    mems = [torch.randn(B, E, s, s + 2) for s in (4, 8, 16)]                 # low -> high resolution
    mask_feature = torch.randn(B, C, 32, 36) * 0.5
    with torch.no_grad():
        out = mod(mems, mask_feature)
        want_masks = out.masks_queries_logits[-1]
        want_cls = torch.nn.functional.linear(out.intermediate_hidden_states[-1].transpose(0, 1), sd["cls_embed.weight"],
                                              sd["cls_embed.bias"])
        cls, mp = rq.stock_mask2former_decoder(sd, mask_feature, mems)
    assert len(out.masks_queries_logits) == n_layers + 1
    torch.testing.assert_close(mp, want_masks, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(cls, want_cls, rtol=2e-3, atol=2e-3)
