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
