"""RSPrompter-query head on the GPU against oracle/restate_query.py:
pixel decoder, Mask2Former decoder + SAM-decoder prompting, instance post-process, full detector."""
import pytest
import torch

pytestmark = pytest.mark.gpu

NQ, NCLS, PTS = 20, 10, 5


def _nerr(a: torch.Tensor, ref: torch.Tensor) -> float:
    return ((a.float().cpu() - ref).norm() / ref.norm()).item()


def _head(nq=NQ):
    from rsprompter_b200 import model_configs, synthetic
    from rsprompter_b200.registry import MODELS
    cfg = model_configs.query_model_cfg("base", NCLS, prompt_shape=(nq, PTS))
    ph = dict(cfg["panoptic_head"])
    ph.update(test_cfg=cfg["test_cfg"])
    head = MODELS.build(ph)
    sd = synthetic.query_head_state_dict(NCLS, nq, PTS, seed=21)
    dec_sd = synthetic.mask_decoder_state_dict(seed=2)
    pe_sd = synthetic.prompt_encoder_state_dict(seed=3)
    full = dict(sd)
    full.update({"mask_decoder.mask_decoder." + k: v for k, v in dec_sd.items()})
    full.update({"sam_" + k: v for k, v in pe_sd.items() if k.startswith("mask_embed.")})
    head.load_state_dict(full, strict=True)
    return head.cuda(), sd, dec_sd, pe_sd


def _feats(B, S, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, 256, S // s, S // s, generator=g).to(torch.bfloat16).float() for s in (4, 8, 16, 32, 64)]


def _nhwc(f):
    return [t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda() for t in f]


def test_groupnorm_topdown_relu():
    import torch.nn.functional as F
    from rsprompter_b200 import _lib
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 128, 24, 16, generator=g).to(torch.bfloat16).float()
    up = torch.randn(2, 128, 12, 8, generator=g).to(torch.bfloat16).float()
    ga, be = 1 + 0.1 * torch.randn(128, generator=g), 0.1 * torch.randn(128, generator=g)
    ref = F.group_norm(x, 32, ga, be, 1e-5)
    for use_up, relu in ((False, False), (True, False), (False, True)):
        r = ref + F.interpolate(up, size=(24, 16), mode="bilinear", align_corners=False) if use_up else ref
        r = F.relu(r) if relu else r
        out = _lib.groupnorm_nhwc(_nhwc([x])[0], ga.cuda(), be.cuda(), 32, up=_nhwc([up])[0] if use_up else None, relu=relu)
        assert (out.float().cpu().permute(0, 3, 1, 2) - r).abs().max().item() < 4e-2


def test_pixel_decoder_matches_oracle():
    from oracle import restate_query
    head, sd, _, _ = _head()
    feats = _feats(2, 256, 11)
    mf_ref, mem_ref = restate_query.pixel_decoder(sd, feats, "pixel_decoder.")
    mf, mems = head.pixel_decoder.forward_nhwc(_nhwc(feats))
    torch.cuda.synchronize()
    assert tuple(mf.shape) == (2, 64, 64, 256) and len(mems) == 3
    for m, r in zip(mems, mem_ref):
        assert _nerr(m.permute(0, 3, 1, 2), r) < 2e-2
    assert _nerr(mf.permute(0, 3, 1, 2), mf_ref) < 2e-2


def _pack_bits(mask: torch.Tensor) -> torch.Tensor:
    """bool [rows, nk] -> int64 words [rows, ceil(nk/64)], bit k%64 of word k/64."""
    rows, nk = mask.shape
    words = (nk + 63) // 64
    m = torch.zeros(rows, words * 64, dtype=torch.int64)
    m[:, :nk] = mask.to(torch.int64)
    m = m.view(rows, words, 64)
    w = torch.zeros(rows, words, dtype=torch.int64)
    for k in range(64):
        w |= m[:, :, k] << k          # bit 63 wraps into the sign bit: same 64-bit pattern
    return w


@pytest.mark.parametrize("nq,nk,masked", [(20, 300, True), (100, 4096, True), (100, 100, False), (130, 64, True)])
def test_mha_small(nq, nk, masked):
    from rsprompter_b200 import _lib
    g = torch.Generator().manual_seed(9)
    B, H, E = 2, 8, 128
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
    out = _lib.mha_small(q.cuda(), k.cuda(), v.cuda(), B, nq, nk, mask=bits)
    assert (out.float().cpu() - ref).abs().max().item() < 2e-2


def test_attn_mask_bits_and_resize():
    import torch.nn.functional as F
    from rsprompter_b200 import _lib
    g = torch.Generator().manual_seed(10)
    for nk in (64, 1000, 4096):
        x = torch.randn(37, nk, generator=g)
        x[3] = -x[3].abs() - 0.1                      # fully masked row -> cleared (M:439-442)
        ref = x < 0
        ref[3] = False
        got = _lib.attn_mask_bits(x.cuda()).cpu()
        exp = _pack_bits(ref)
        if nk % 64:
            tail = ~((1 << (nk % 64)) - 1)          # bits past nk are "masked" unless the row was cleared
            exp[:, -1] |= tail
            exp[3, -1] = 0
        assert torch.equal(got, exp)
    x = torch.randn(2, 64, 48, 16, generator=g).to(torch.bfloat16)
    for hw in ((16, 12), (8, 6), (4, 3), (24, 20)):
        ref = F.interpolate(x.float().permute(0, 3, 1, 2), size=hw, mode="bilinear", align_corners=False)
        out = _lib.resize_bilinear_nhwc(x.cuda(), hw).float().cpu().permute(0, 3, 1, 2)
        assert (out - ref).abs().max().item() < 2e-2


def test_query_head_matches_oracle():
    """cls / mask_pred_plus / SAM-decoder masks of the last layer (M:395-463) at 512^2, 2 images x 20 queries."""
    from oracle import restate, restate_query
    from rsprompter_b200 import sam_config
    head, sd, dec_sd, pe_sd = _head()
    B, S = 2, 512
    feats = _feats(B, S, 13)
    g = torch.Generator().manual_seed(14)
    h = S // 16
    emb = torch.randn(B, 256, h, h, generator=g)
    pos = torch.randn(1, 256, h, h, generator=g).repeat(B, 1, 1, 1)
    cls_ref, mp_ref, mpp_ref = restate_query.mask2former_head(sd, sam_config.decoder_arch(), dec_sd, pe_sd, feats, emb, pos,
                                                              points=PTS)
    emb_rows = emb.permute(0, 2, 3, 1).reshape(B * h * h, 256).contiguous().cuda()
    pos_rows = pos[0].permute(1, 2, 0).reshape(h * h, 256).contiguous().cuda()
    cls, masks, mpp = head.forward_nhwc(_nhwc(feats), emb_rows, pos_rows, (h, h))
    torch.cuda.synchronize()
    assert tuple(masks.shape) == (B * NQ, 4 * h, 4 * h) and tuple(mpp.shape) == (B, NQ, S // 4, S // 4)
    assert _nerr(mpp, mpp_ref) < 3e-2
    assert _nerr(cls, cls_ref) < 3e-2
    assert _nerr(masks.view(B, NQ, 4 * h, 4 * h), mp_ref) < 3e-2


def test_instance_postprocess_matches_oracle():
    from oracle import restate_query
    import torch.nn.functional as F
    from rsprompter_b200.registry import MODELS
    g = torch.Generator().manual_seed(17)
    B, nq, hm, S, K = 2, 20, 64, 256, 15
    fh = MODELS.build(dict(type="RSMaskFormerFusionHead", num_things_classes=NCLS, num_stuff_classes=0,
                           test_cfg=dict(max_per_image=K, instance_on=True, panoptic_on=False)))
    cls = torch.randn(B, nq, NCLS + 1, generator=g) * 2
    yy, xx = torch.meshgrid(torch.arange(hm), torch.arange(hm), indexing="ij")
    cy, cx, r = (torch.rand(B * nq, 1, 1, generator=g) * hm for _ in range(3))
    logit = (r * 0.4 + 3 - ((yy - cy) ** 2 + (xx - cx) ** 2).sqrt()) * 1.7 + 0.05 * torch.randn(B * nq, hm, hm, generator=g)
    out = fh.instance_postprocess_batched(cls.cuda(), logit.cuda().contiguous(), (S, S))
    torch.cuda.synchronize()
    up = F.interpolate(logit.view(B, nq, hm, hm), size=(S, S), mode="bilinear", align_corners=False)
    for b in range(B):
        ref = restate_query.instance_postprocess(cls[b], up[b], NCLS, K)
        key = lambda q, l: (q * NCLS + l).tolist()  # noqa: E731
        order_ref = {k: i for i, k in enumerate(key(ref["query"], ref["labels"]))}
        got = key(out["query"][b].cpu(), out["labels"][b].cpu())
        assert sorted(got) == sorted(order_ref)
        idx = torch.tensor([order_ref[k] for k in got])
        m_ref, m = ref["masks"][idx], out["masks"][b].cpu()
        near = (up[b][ref["query"][idx]].abs() < 1e-4)
        assert ((m != m_ref) & ~near).sum().item() == 0
        assert torch.allclose(out["scores"][b].cpu(), ref["scores"][idx], rtol=2e-3, atol=1e-5)
        clean = ~((m != m_ref).flatten(1).any(1))
        assert torch.equal(out["bboxes"][b].cpu()[clean], ref["bboxes"][idx][clean])


def test_query_detector_predict():
    from rsprompter_b200 import model_configs, sam_config, synthetic
    from rsprompter_b200.registry import MODELS
    cfg = model_configs.query_model_cfg("base", NCLS, prompt_shape=(30, PTS), mmpretrain_img_size=512)
    model = MODELS.build(cfg)
    arch = model.backbone.vision_encoder.arch
    model.load_state_dict(synthetic.query_detector_state_dict(arch, NCLS, 0, nq=30, seed=3, pseudo_neck=True), strict=True)
    model = model.cuda()
    torch.manual_seed(3)
    x = torch.randn(2, 3, 512, 512).cuda()
    res = model.predict(x)
    torch.cuda.synchronize()
    assert len(res) == 2
    for ds in res:
        p = ds.pred_instances
        assert p.masks.shape == (30, 512, 512) and p.masks.dtype == torch.bool
        assert p.bboxes.shape == (30, 4) and p.scores.shape == (30,) and p.labels.max().item() < NCLS
        assert torch.isfinite(p.scores).all()
        # tight boxes of the returned masks (mask2bbox, mask/utils.py:56-77)
        for i in range(0, 30, 7):
            ys, xs = torch.nonzero(p.masks[i], as_tuple=True)
            if ys.numel():
                exp = torch.tensor([xs.min(), ys.min(), xs.max() + 1, ys.max() + 1], dtype=torch.float32)
                assert torch.equal(p.bboxes[i].cpu(), exp)
            else:
                assert p.bboxes[i].abs().sum().item() == 0


def test_instance_postprocess_rescale_matches_oracle():
    """Resized + padded image (M:652-656, 679-691): logits -> batch shape -> crop -> ori_shape, then mask / score / box."""
    from oracle import restate_query
    from rsprompter_b200.registry import MODELS
    g = torch.Generator().manual_seed(19)
    B, nq, hm, K = 2, 12, 64, 9
    batch, ori = (256, 256), (150, 210)
    s = min(batch[0] / ori[0], batch[1] / ori[1])
    new_hw = (int(ori[0] * s + 0.5), int(ori[1] * s + 0.5))
    meta = dict(ori_shape=ori, batch_input_shape=batch, scale_factor=(new_hw[1] / ori[1], new_hw[0] / ori[0]))
    sf = meta["scale_factor"]
    metas = [None, dict(ori_hw=ori, crop_hw=(min(int(ori[0] * sf[1]), batch[0]), min(int(ori[1] * sf[0]), batch[1])), scale_factor=sf)]
    fh = MODELS.build(dict(type="RSMaskFormerFusionHead", num_things_classes=NCLS, num_stuff_classes=0,
                           test_cfg=dict(max_per_image=K, instance_on=True, panoptic_on=False)))
    cls = torch.randn(B, nq, NCLS + 1, generator=g) * 2
    yy, xx = torch.meshgrid(torch.arange(hm), torch.arange(hm), indexing="ij")
    cy, cx, r = (torch.rand(B * nq, 1, 1, generator=g) * hm for _ in range(3))
    logit = (r * 0.4 + 3 - ((yy - cy) ** 2 + (xx - cx) ** 2).sqrt()) * 1.7 + 0.05 * torch.randn(B * nq, hm, hm, generator=g)
    out = fh.instance_postprocess_batched(cls.cuda(), logit.cuda().contiguous(), batch, metas=metas, rescale=True)
    torch.cuda.synchronize()
    up = restate_query.fusion_rescale(logit.view(B, nq, hm, hm)[1], meta)            # [nq, 150, 210]
    ref = restate_query.instance_postprocess(cls[1], up, NCLS, K)
    key = lambda q, l: (q * NCLS + l).tolist()  # noqa: E731
    order_ref = {k: i for i, k in enumerate(key(ref["query"], ref["labels"]))}
    got = key(out["query"][1].cpu(), out["labels"][1].cpu())
    assert sorted(got) == sorted(order_ref)
    idx = torch.tensor([order_ref[k] for k in got])
    m = out["masks"][1].cpu()
    assert m.shape == (K, *ori) and out["masks"][0].shape == (K, *batch)
    m_ref = ref["masks"][idx]
    near = up[ref["query"][idx]].abs() < 1e-4
    assert ((m != m_ref) & ~near).sum().item() == 0
    assert torch.allclose(out["scores"][1].cpu(), ref["scores"][idx], rtol=2e-3, atol=1e-5)
    clean = ~((m != m_ref).flatten(1).any(1))
    assert torch.equal(out["bboxes"][1].cpu()[clean], ref["bboxes"][idx][clean])
