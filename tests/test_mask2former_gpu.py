"""SAMSegMask2Former (M:1247-1274 over mmdet Mask2Former): the stock Mask2FormerHead at feat_channels
256 (8 heads x 32) on the GPU against oracle/restate_query.py (stock_mask2former_head / samseg_mask2former_predict):
the 256-channel instantiations of GroupNorm, MSDeformAttn and the small MHA, the head, the whole detector."""
import pytest
import torch

pytestmark = pytest.mark.gpu

NQ, NCLS = 20, 10


def _nerr(a: torch.Tensor, ref: torch.Tensor) -> float:
    return ((a.float().cpu() - ref).norm() / ref.norm()).item()


def _nhwc(f):
    return [t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda() for t in f]


def _feats(B, S, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, 256, S // s, S // s, generator=g).to(torch.bfloat16).float() for s in (4, 8, 16, 32, 64)]


def _pack_bits(mask: torch.Tensor) -> torch.Tensor:
    rows, nk = mask.shape
    words = (nk + 63) // 64
    m = torch.zeros(rows, words * 64, dtype=torch.int64)
    m[:, :nk] = mask.to(torch.int64)
    m = m.view(rows, words, 64)
    w = torch.zeros(rows, words, dtype=torch.int64)
    for k in range(64):
        w |= m[:, :, k] << k
    return w


def test_groupnorm_8_channels_per_group():
    import torch.nn.functional as F
    from rsprompter_b200 import _lib
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(2, 256, 24, 16, generator=g) * 2 + 3).to(torch.bfloat16).float()
    up = torch.randn(2, 256, 12, 8, generator=g).to(torch.bfloat16).float()
    ga, be = 1 + 0.1 * torch.randn(256, generator=g), 0.1 * torch.randn(256, generator=g)
    ref = F.group_norm(x, 32, ga, be, 1e-5)
    for use_up, relu in ((False, False), (True, False), (False, True)):
        r = ref + F.interpolate(up, size=(24, 16), mode="bilinear", align_corners=False) if use_up else ref
        r = F.relu(r) if relu else r
        out = _lib.groupnorm_nhwc(_nhwc([x])[0], ga.cuda(), be.cuda(), 32, up=_nhwc([up])[0] if use_up else None, relu=relu)
        assert (out.float().cpu().permute(0, 3, 1, 2) - r).abs().max().item() < 4e-2


@pytest.mark.parametrize("nq,nk,masked", [(20, 300, True), (100, 4096, True), (100, 100, False), (130, 64, True)])
def test_mha_small_head_dim_32(nq, nk, masked):
    from rsprompter_b200 import _lib
    g = torch.Generator().manual_seed(9)
    B, H, E = 2, 8, 256
    q, k, v = (torch.randn(B * n, E, generator=g).to(torch.bfloat16) for n in (nq, nk, nk))
    sp = lambda t, n: t.float().view(B, n, H, E // H).transpose(1, 2)  # noqa: E731
    s = (sp(q, nq) @ sp(k, nk).transpose(-1, -2)) * (E // H) ** -0.5
    bits = None
    if masked:
        mask = torch.rand(B, nq, nk, generator=g) < 0.6
        mask[0, 3] = False
        mask[1, 5, : nk - 1] = True
        mask[1, 5, nk - 1] = False
        s = s.masked_fill(mask[:, None], float("-inf"))
        bits = _pack_bits(mask.view(B * nq, nk)).cuda()
    ref = (s.softmax(-1) @ sp(v, nk)).transpose(1, 2).reshape(B * nq, E)
    qkv = torch.cat([q[: B * min(nq, nk)], k[: B * min(nq, nk)]], dim=1) if nq == nk else None
    out = _lib.mha_small(q.cuda(), k.cuda(), v.cuda(), B, nq, nk, mask=bits, head_dim=32)
    torch.cuda.synchronize()
    assert out.shape == (B * nq, E)
    assert (out.float().cpu() - ref).abs().max().item() < 2e-2
    if qkv is not None:      # strided row views (the fused [Q | K] projection of the self-attention)
        qk = qkv.cuda()
        out2 = _lib.mha_small(qk[:, :E], qk[:, E:], v.cuda(), B, nq, nk, mask=bits, head_dim=32)
        assert torch.equal(out2, out)


def _head(nq=NQ, seed=31, dec_layers=9):
    from rsprompter_b200 import model_configs, synthetic
    from rsprompter_b200.registry import MODELS
    cfg = model_configs.mask2former_model_cfg("base", NCLS, num_queries=nq)
    ph = dict(cfg["panoptic_head"])
    ph.update(test_cfg=cfg["test_cfg"])
    ph["transformer_decoder"] = dict(ph["transformer_decoder"], num_layers=dec_layers)
    head = MODELS.build(ph)
    sd = synthetic.mask2former_head_state_dict(NCLS, nq, dec_layers=dec_layers, seed=seed)
    head.load_state_dict(sd, strict=True)
    return head.cuda(), sd


def test_pixel_decoder_256_matches_oracle():
    """MSDeformAttnPixelDecoder at feat_channels 256: input convs + GN(32 groups of 8), 3 deformable encoder layers
    (8 heads x 32), FPN top-down, mask_feature."""
    from oracle import restate_query
    head, sd = _head()
    feats = _feats(2, 256, 11)
    mf_ref, mem_ref = restate_query.pixel_decoder(sd, feats, "pixel_decoder.")
    mf, mems = head.pixel_decoder.forward_nhwc(_nhwc(feats))
    torch.cuda.synchronize()
    assert tuple(mf.shape) == (2, 64, 64, 256) and len(mems) == 3 and mems[0].shape[-1] == 256
    for m, r in zip(mems, mem_ref):
        assert _nerr(m.permute(0, 3, 1, 2), r) < 2e-2
    assert _nerr(mf.permute(0, 3, 1, 2), mf_ref) < 2e-2


@pytest.mark.parametrize("dec_layers,tol", [(3, 5e-2), (9, 7e-2)])
def test_stock_mask2former_head_matches_oracle(dec_layers, tol):
    """cls / mask logits of the last decoder layer (mask2former_head.py:382-460) at 512^2, 2 x 20 queries, on random
    features.  Every layer thresholds its mask logits into the next layer's attention mask, so bf16-level differences
    on pixels at the threshold flip mask bits and compound with depth: at 3 layers the masks measured < 3e-2 and the
    class logits 4.2e-2 (5e-2 asserted), the shipped 9 layers 4.6e-2 on this input (7e-2 asserted); the whole detector
    on encoder features is pinned in test_samseg_mask2former_end_to_end_matches_oracle."""
    from oracle import restate_query
    head, sd = _head(dec_layers=dec_layers)
    B, S = 2, 512
    feats = _feats(B, S, 13)
    with torch.no_grad():
        cls_ref, mp_ref = restate_query.stock_mask2former_head(sd, feats)
    cls, masks = head.forward_nhwc(_nhwc(feats))
    torch.cuda.synchronize()
    assert tuple(cls.shape) == (B, NQ, NCLS + 1) and tuple(masks.shape) == (B * NQ, S // 4, S // 4)
    assert _nerr(masks.view(B, NQ, S // 4, S // 4), mp_ref) < tol
    assert _nerr(cls, cls_ref) < tol


def test_samseg_mask2former_end_to_end_matches_oracle():
    """ViT-B 1024^2, bs 1, 100 queries: instance keys, scores and the last layer's mask logits against the fp32 oracle
    (same reading of the tolerances as test_e2e_gpu.test_query_1024_end_to_end_matches_oracle: thresholded masks feed
    the next layer's attention masks, so the worst query is not a rounding measure - mean / p99.9 are asserted)."""
    import json
    import os
    from oracle import restate_query as rq
    from rsprompter_b200 import model_configs, sam_config, synthetic
    from rsprompter_b200.model_configs import SELECT_LAYERS
    from rsprompter_b200.registry import MODELS
    nq, size = 100, 1024
    m = MODELS.build(model_configs.mask2former_model_cfg("base", NCLS, num_queries=nq))
    arch = sam_config.VISION_ARCHS["base"]
    sel = SELECT_LAYERS["base"]
    sd = synthetic.mask2former_detector_state_dict(arch, NCLS, len(sel), nq=nq, seed=6)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    torch.manual_seed(6)
    x = torch.randn(1, 3, size, size)
    with torch.no_grad():
        ref = rq.samseg_mask2former_predict(sd, arch, x, NCLS, sel, max_per_image=nq)[0]
    out = m.predict(x.cuda())[0].pred_instances
    raw = m.predict_raw(x.cuda())
    res = m.panoptic_fusion_head.instance_postprocess_batched(raw["cls"], raw["mask_logits"], (size, size))
    rec = m.predict_records(x.cuda()).instances()[0]
    torch.cuda.synchronize()
    assert len(out) == nq and out.masks.shape == (nq, size, size) and out.masks.dtype == torch.bool
    assert torch.equal(rec["masks"], out.masks) and torch.equal(rec["bboxes"], out.bboxes)
    dl = (raw["mask_logits"].cpu() - ref["mask_logits"]).abs()
    dc = (raw["cls"][0].cpu() - ref["cls"]).abs()
    key = lambda q, l: (q * NCLS + l).tolist()  # noqa: E731
    kr = {k: i for i, k in enumerate(key(ref["query"], ref["labels"]))}
    kg = key(res["query"][0].cpu(), res["labels"][0].cpu())
    shared = [(i, kr[k]) for i, k in enumerate(kg) if k in kr]
    gi, ri = torch.tensor([p[0] for p in shared]), torch.tensor([p[1] for p in shared])
    gm, rm = res["masks"][0].cpu()[gi], ref["masks"][ri]
    scale = ref["mask_logits"].abs().max().item()
    rep = dict(keys_gpu=len(kg), keys_ref=len(kr), shared=len(shared), logit_max_diff=dl.max().item(),
               logit_mean_diff=dl.mean().item(), logit_scale=scale,
               logit_p999_diff=dl.flatten().kthvalue(int(0.999 * dl.numel())).values.item(),
               cls_max_diff=dc.max().item(), cls_scale=ref["cls"].abs().max().item(),
               score_max_diff=(res["scores"][0].cpu()[gi] - ref["scores"][ri]).abs().max().item(),
               mask_disagree=(gm != rm).float().mean().item())
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "parity_e2e_mask2former_vitb_1024.json"), "w") as f:
            json.dump(rep, f, indent=1)
    except OSError:
        pass
    print("mask2former e2e", rep)
    tol = 2e-2 * max(1.0, scale)
    assert rep["shared"] >= 0.9 * nq
    assert rep["logit_mean_diff"] <= tol / 2 and rep["logit_p999_diff"] <= 5 * tol
    assert rep["cls_max_diff"] <= 5 * 2e-2 * max(1.0, rep["cls_scale"])
    assert rep["mask_disagree"] <= 2e-2
    assert rep["score_max_diff"] <= 5e-2
