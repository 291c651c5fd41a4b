"""SAMSegMaskRCNN (M:1218-1244 over mmdet MaskRCNN): the stock StandardRoIHead / FCNMaskHead mask branch
on the GPU against the oracle (oracle.restate_anchor.fcn_mask_head / paste_masks_in_boxes / maskrcnn_predict) and
against the fixture generated from the reference's own fcn_mask_head.py (tests/golden/reference_maskrcnn.pt)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

NUM_CLASSES = 10
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_maskrcnn.pt")


def test_mask_paste_boxes_matches_reference_fixture():
    from rsprompter_b200 import _lib
    f = torch.load(GOLDEN, weights_only=False)["do_paste_mask"]
    H, W = f["hw"]
    got = _lib.mask_paste_boxes(f["probs"][:, 0].contiguous().cuda(), f["boxes"].cuda(), (H, W), 0.5)
    torch.cuda.synchronize()
    ref = f["out"] >= 0.5
    near = (f["out"] - 0.5).abs() < 1e-5
    assert got.shape == ref.shape and got.dtype == torch.bool
    assert ((got.cpu() != ref) & ~near).sum().item() == 0
    assert (got.cpu() != ref).sum().item() <= 2


@pytest.mark.parametrize("H,W", [(1024, 1024), (333, 500)])
def test_mask_paste_boxes_matches_oracle_on_full_canvas(H, W):
    """Full-size canvases (W % 16 == 0: 16-byte stores; any other W: byte stores), boxes partly outside, degenerate
    boxes; bit-equal to the oracle away from fp ties and zero outside the box's 1-pixel bilinear fringe."""
    from oracle import restate_anchor as ra
    from rsprompter_b200 import _lib
    g = torch.Generator().manual_seed(H + W)
    n = 12
    probs = torch.rand(n, 28, 28, generator=g)
    boxes = torch.rand(n, 4, generator=g) * torch.tensor([W * 0.7, H * 0.7, W * 0.7, H * 0.7])
    boxes[:, 2:] = boxes[:, :2] + torch.rand(n, 2, generator=g) * torch.tensor([W * 0.4, H * 0.4]) + 1.0
    boxes[0] = torch.tensor([-20.0, -10.0, W + 15.0, H + 5.0])
    boxes[1] = torch.tensor([W / 2, 10.0, W / 2, 80.0])
    boxes[2] = torch.tensor([0.0, 0.0, 0.0, 0.0])
    got = _lib.mask_paste_boxes(probs.cuda(), boxes.cuda(), (H, W), 0.5).cpu()
    val = ra.paste_masks_in_boxes(probs, boxes, H, W)
    ref = val >= 0.5
    near = (val - 0.5).abs() < 1e-5
    assert ((got != ref) & ~near).sum().item() == 0
    ys, xs = torch.arange(H).view(1, H, 1) + 0.5, torch.arange(W).view(1, 1, W) + 0.5
    for i in range(3, n):
        b = boxes[i]
        bw, bh = (b[2] - b[0]) / 28, (b[3] - b[1]) / 28          # one RoI cell: the reach of the zero-padded fringe
        outside = (xs < b[0] - bw) | (xs > b[2] + bw) | (ys < b[1] - bh) | (ys > b[3] + bh)
        assert not (got[i:i + 1] & outside).any()


def _dump(name: str, obj) -> None:
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", name), "w") as f:
            json.dump(obj, f, indent=1)
    except OSError:
        pass


def _match_boxes(gb, gl, rb, rl, thr=0.9):
    """greedy one-to-one matching of GPU detections to oracle detections: same label, IoU >= thr."""
    if gb.numel() == 0 or rb.numel() == 0:
        return []
    lt = torch.maximum(gb[:, None, :2], rb[None, :, :2])
    br = torch.minimum(gb[:, None, 2:], rb[None, :, 2:])
    inter = (br - lt).clamp(min=0).prod(-1)
    area = lambda x: (x[:, 2] - x[:, 0]).clamp(min=0) * (x[:, 3] - x[:, 1]).clamp(min=0)  # noqa: E731
    iou = inter / (area(gb)[:, None] + area(rb)[None, :] - inter + 1e-9)
    iou[gl[:, None] != rl[None, :]] = 0
    pairs, used = [], set()
    for i in iou.max(dim=1).values.argsort(descending=True).tolist():
        j = int(iou[i].argmax())
        if iou[i, j] >= thr and j not in used:
            used.add(j)
            pairs.append((i, j))
    return pairs


def _head(seed=5):
    from rsprompter_b200 import synthetic
    from rsprompter_b200.registry import MODELS
    head = MODELS.build(dict(type="FCNMaskHead", num_convs=4, in_channels=256, conv_out_channels=256,
                             num_classes=NUM_CLASSES))
    sd = synthetic.fcn_mask_head_state_dict(NUM_CLASSES, seed=seed)
    head.load_state_dict(sd, strict=True)
    return head.cuda(), sd


def test_fcn_mask_head_matches_oracle():
    from oracle import restate_anchor as ra
    head, sd = _head()
    g = torch.Generator().manual_seed(5)
    n = 37
    feats = torch.randn(n, 14, 14, 256, generator=g).to(torch.bfloat16)
    labels = torch.randint(0, NUM_CLASSES, (n,), generator=g)
    out = head.forward_rows(feats.reshape(n, -1).cuda())
    sel = head.select(out, labels.cuda())
    torch.cuda.synchronize()
    assert out.shape[:3] == (n, 28, 28) and out.shape[3] >= NUM_CLASSES and sel.shape == (n, 28, 28)
    with torch.no_grad():
        ref = ra.fcn_mask_head(sd, feats.float().permute(0, 3, 1, 2), prefix="")           # [n, C, 28, 28]
    got = out[..., :NUM_CLASSES].permute(0, 3, 1, 2).cpu()
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= 2e-2 * max(1.0, scale)
    assert torch.equal(sel.cpu(), got[torch.arange(n), labels])


def test_samseg_maskrcnn_end_to_end_matches_oracle():
    """ViT-B 1024^2, bs 1: detections matched to the fp32 oracle's (label + IoU), scores close, the 28x28 mask logits on
    the GPU path's own boxes within the bf16 tolerance, pasted masks equal away from the 0.5 boundary."""
    from oracle import restate_anchor as ra
    from rsprompter_b200 import model_configs, synthetic
    from rsprompter_b200.model_configs import SELECT_LAYERS
    from rsprompter_b200.registry import MODELS
    size, seed = 1024, 9
    m = MODELS.build(model_configs.maskrcnn_model_cfg("base", NUM_CLASSES))
    arch = m.backbone.vision_encoder.arch
    sel = SELECT_LAYERS["base"]
    sd = synthetic.maskrcnn_detector_state_dict(arch, NUM_CLASSES, len(sel), seed=seed)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    torch.manual_seed(seed)
    x = torch.randn(1, 3, size, size)
    out = m.predict(x.cuda())[0].pred_instances
    cap: dict = {}
    raw = m.predict_raw(x.cuda(), capture=cap)
    torch.cuda.synchronize()
    n = int(raw["counts"][0])
    gb, gs, gl = out.bboxes.cpu(), out.scores.cpu(), out.labels.cpu()
    assert n > 0 and len(out) == n and out.masks.shape == (n, size, size) and out.masks.dtype == torch.bool
    with torch.no_grad():
        ref = ra.maskrcnn_predict(sd, arch, x, NUM_CLASSES, sel, extra_boxes=[gb])[0]
    pairs = _match_boxes(gb, gl, ref["bboxes"], ref["labels"])
    gi, ri = torch.tensor([p[0] for p in pairs]), torch.tensor([p[1] for p in pairs])
    xl = ref["extra_mask_logits"]                                       # [n, C, 28, 28] on the GPU path's boxes
    glog = cap["mask_logits_all"][:n, :, :, :NUM_CLASSES].permute(0, 3, 1, 2).cpu()
    dl = (glog - xl).abs()
    xsel = torch.sigmoid(xl[torch.arange(n), gl])
    val = ra.paste_masks_in_boxes(xsel, gb, size, size)
    near = (val - 0.5).abs() < 5e-3
    dm = out.masks.cpu() != (val >= 0.5)
    rep = dict(n_gpu=n, n_ref=int(ref["bboxes"].shape[0]), matched=len(pairs),
               score_max_diff=(gs[gi] - ref["scores"][ri]).abs().max().item() if pairs else None,
               score_mean_diff=(gs[gi] - ref["scores"][ri]).abs().mean().item() if pairs else None,
               box_median_diff=(gb[gi] - ref["bboxes"][ri]).abs().amax(dim=1).median().item() if pairs else None,
               logit_max_diff=dl.max().item(), logit_mean_diff=dl.mean().item(), logit_scale=xl.abs().max().item(),
               mask_disagree=dm.float().mean().item(), mask_disagree_off_boundary=(dm & ~near).float().mean().item())
    _dump("parity_e2e_maskrcnn_vitb_1024.json", rep)
    print("maskrcnn e2e", rep)
    assert rep["matched"] >= 0.8 * max(rep["n_gpu"], rep["n_ref"])
    # softmax scores of matched detections: the RoIs differ by a fraction of a pixel (bf16 RPN / neck), and the synthetic
    # fc_cls is sharp, so a single score near 0.5 moves by up to ~0.1 (measured 0.09, mean 0.01); the class logits on
    # IDENTICAL RoI features are pinned to 2e-2 in test_anchor_gpu (Shared2FCBBoxHead)
    assert rep["score_max_diff"] <= 0.15 and rep["score_mean_diff"] <= 2.5e-2
    assert rep["logit_max_diff"] <= 2e-2 * max(1.0, rep["logit_scale"])
    assert rep["mask_disagree_off_boundary"] <= 2e-4


def test_samseg_maskrcnn_rescale_and_record():
    """Resized image (scale_factor != 1): boxes / masks in the original frame, the same paste the oracle's
    fcn_mask_predict_single does; predict_records() == predict() on the batch-shaped path."""
    from oracle import restate_anchor as ra
    from rsprompter_b200 import model_configs, synthetic
    from rsprompter_b200.registry import MODELS, make_data_samples
    m = MODELS.build(model_configs.maskrcnn_model_cfg("base", NUM_CLASSES))
    arch = m.backbone.vision_encoder.arch
    m.load_state_dict(synthetic.maskrcnn_detector_state_dict(arch, NUM_CLASSES, 6, seed=11), strict=True)
    m = m.cuda()
    torch.manual_seed(11)
    x = torch.randn(1, 3, 1024, 1024).cuda()
    base = m.predict(x)[0].pred_instances
    rec = m.predict_records(x).instances()[0]
    assert torch.equal(rec["bboxes"], base.bboxes) and torch.equal(rec["masks"], base.masks)
    assert torch.equal(rec["labels"], base.labels) and torch.equal(rec["scores"], base.scores)
    ds = make_data_samples(1, (1024, 1024))
    ds[0].set_metainfo(dict(ori_shape=(700, 811), img_shape=(884, 1024), scale_factor=(1024 / 811, 884 / 700)))
    cap: dict = {}
    raw = m.predict_raw(m._attach_img_shapes(ds, x), capture=cap)          # boxes clipped to img_shape, as predict() does
    out = m.predict(x, ds, rescale=True)[0].pred_instances
    torch.cuda.synchronize()
    n = len(out)
    assert n == int(raw["counts"][0]) and out.masks.shape == (n, 700, 811)
    assert raw["bboxes"][0, :n, 1::2].max().item() <= 884
    logits = cap["mask_logits_all"][:n, :, :, :NUM_CLASSES].permute(0, 3, 1, 2).cpu()
    masks, boxes = ra.fcn_mask_predict_single(logits, raw["bboxes"][0, :n].cpu(), out.labels.cpu(), (700, 811),
                                              (1024 / 811, 884 / 700), rescale=True)
    torch.testing.assert_close(out.bboxes.cpu(), boxes, rtol=0, atol=1e-3)
    assert (out.masks.cpu() != masks).float().mean().item() <= 1e-4


def test_zero_border_nhwc():
    from rsprompter_b200 import _lib
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 16, 16, 64, generator=g).to(torch.bfloat16).cuda()
    ref = x.clone()
    ref[:, 0] = 0
    ref[:, -1] = 0
    ref[:, :, 0] = 0
    ref[:, :, -1] = 0
    _lib.zero_border_nhwc(x)
    assert torch.equal(x, ref) and x[:, 1:-1, 1:-1].abs().sum() > 0


def test_fcn_mask_head_canvas_path_equals_im2col_path():
    """The 16x16-canvas implicit-GEMM path against the same head run through the im2col path (31 RoI channels do not
    qualify for the canvas, so a 256-channel head is forced through both by toggling the geometry check)."""
    from rsprompter_b200 import _lib
    head, _ = _head()
    g = torch.Generator().manual_seed(8)
    feats = torch.randn(9, 14 * 14 * 256, generator=g).to(torch.bfloat16).cuda()
    a = head.forward_rows(feats)
    ok = _lib.conv3x3_ok
    try:
        _lib.conv3x3_ok = lambda *args: False
        b = head.forward_rows(feats)
    finally:
        _lib.conv3x3_ok = ok
    torch.cuda.synchronize()
    assert a.shape == b.shape and not a.is_contiguous() and b.is_contiguous()
    assert (a - b).abs().max().item() <= 2e-2 * max(1.0, b.abs().max().item())
