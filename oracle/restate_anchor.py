"""fp32 CPU restatement of the RSPrompter-anchor head path (TEST INFRASTRUCTURE, see
oracle/__init__.py): RSFPN neck -> RPNHead -> RSPrompterAnchorRoIPromptHead -> mask head ->
SAM decoder -> mask post-processing.  State-dict keys follow the reference module tree
(``neck.feature_aggregator.downconvs.0.0.weight`` ...).

Third-party bricks absent from /root/reference are replaced by their documented equivalents
(unverifiable offline, flagged in SURVEY.md 8c): mmcv.ops.nms -> torchvision.ops.nms (same IoU
formula, suppress when IoU > thr, offset 0), mmcv.ops.RoIAlign(aligned=True, sampling_ratio=0,
pool_mode='avg') -> torchvision.ops.roi_align(aligned=True, sampling_ratio=0).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import restate

try:
    import torchvision.ops as tvops
except Exception:  # noqa: BLE001
    tvops = None


# --------------------------------------------------------------------------------------------
# necks
# --------------------------------------------------------------------------------------------
def _bn_eval(sd: dict, p: str, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training=False, eps=eps)


def _conv_bn_relu(sd: dict, pc: str, pb: str, x: torch.Tensor, padding: int = 0, stride: int = 1) -> torch.Tensor:
    x = F.conv2d(x, sd[pc + ".weight"], sd.get(pc + ".bias"), padding=padding, stride=stride)
    return F.relu(_bn_eval(sd, pb, x))


def feature_aggregator(sd: dict, hidden_states, select_layers, prefix: str = "") -> torch.Tensor:
    """RSFeatureAggregator.forward (M:1042-1057); BatchNorm in eval mode."""
    p = prefix
    inputs = [h.permute(0, 3, 1, 2) for h in hidden_states]
    feats = []
    for idx, il in enumerate(select_layers):
        x = _conv_bn_relu(sd, f"{p}downconvs.{idx}.0", f"{p}downconvs.{idx}.1", inputs[il])
        x = _conv_bn_relu(sd, f"{p}downconvs.{idx}.3", f"{p}downconvs.{idx}.4", x, padding=1)
        feats.append(x)
    x = None
    for idx, hs in enumerate(feats):
        if x is not None:
            hs = x + hs
        res = _conv_bn_relu(sd, f"{p}hidden_convs.{idx}.0", f"{p}hidden_convs.{idx}.1", hs, padding=1)
        x = hs + res
    x = _conv_bn_relu(sd, f"{p}fusion_conv.0", f"{p}fusion_conv.1", x)
    x = _conv_bn_relu(sd, f"{p}fusion_conv.3", f"{p}fusion_conv.4", x, padding=1)
    return F.conv2d(x, sd[f"{p}fusion_conv.6.weight"], sd[f"{p}fusion_conv.6.bias"], padding=1)


def pseudo_feature_aggregator(sd: dict, x: torch.Tensor, prefix: str = "") -> torch.Tensor:
    """PseudoFeatureAggregator.forward (M:980-984): conv1x1, LN2d, conv3x3, LN2d, conv3x3, LN2d."""
    p = prefix + "channel_fusion."
    ln = restate.layer_norm_channels_first
    x = ln(F.conv2d(x, sd[p + "0.weight"]), sd[p + "1.weight"], sd[p + "1.bias"], 1e-6)
    x = ln(F.conv2d(x, sd[p + "2.weight"], padding=1), sd[p + "3.weight"], sd[p + "3.bias"], 1e-6)
    x = ln(F.conv2d(x, sd[p + "4.weight"], padding=1), sd[p + "5.weight"], sd[p + "5.bias"], 1e-6)
    return x


def simple_fpn(sd: dict, x: torch.Tensor, prefix: str = "", num_outs: int = 5, norm_key: str = "ln"):
    """RSSimpleFPN.forward (M:1334-1363).  ConvModule(norm_cfg=LN2d) = bias-free conv + LN2d."""
    p = prefix
    ln = restate.layer_norm_channels_first
    f1 = F.conv_transpose2d(x, sd[p + "fpn1.0.weight"], sd[p + "fpn1.0.bias"], stride=2)
    f1 = F.gelu(ln(f1, sd[p + "fpn1.1.weight"], sd[p + "fpn1.1.bias"], 1e-6))
    f1 = F.conv_transpose2d(f1, sd[p + "fpn1.3.weight"], sd[p + "fpn1.3.bias"], stride=2)
    f2 = F.conv_transpose2d(x, sd[p + "fpn2.0.weight"], sd[p + "fpn2.0.bias"], stride=2)
    f3 = x
    f4 = F.max_pool2d(x, 2, 2)
    outs = []
    for i, f in enumerate([f1, f2, f3, f4]):
        lat = F.conv2d(f, sd[f"{p}lateral_convs.{i}.conv.weight"])
        lat = ln(lat, sd[f"{p}lateral_convs.{i}.{norm_key}.weight"], sd[f"{p}lateral_convs.{i}.{norm_key}.bias"], 1e-6)
        o = F.conv2d(lat, sd[f"{p}fpn_convs.{i}.conv.weight"], padding=1)
        o = ln(o, sd[f"{p}fpn_convs.{i}.{norm_key}.weight"], sd[f"{p}fpn_convs.{i}.{norm_key}.bias"], 1e-6)
        outs.append(o)
    for _ in range(num_outs - 4):
        outs.append(F.max_pool2d(outs[-1], 1, stride=2))
    return outs


# --------------------------------------------------------------------------------------------
# RPN
# --------------------------------------------------------------------------------------------
def base_anchors(base_size: float, scales, ratios) -> torch.Tensor:
    """AnchorGenerator.gen_single_level_base_anchors (anchor_generator.py:161-205), scale_major,
    center_offset 0."""
    scales = torch.tensor(scales, dtype=torch.float32)
    ratios = torch.tensor(ratios, dtype=torch.float32)
    hr = torch.sqrt(ratios)
    wr = 1 / hr
    ws = (base_size * wr[:, None] * scales[None, :]).view(-1)
    hs = (base_size * hr[:, None] * scales[None, :]).view(-1)
    return torch.stack([-0.5 * ws, -0.5 * hs, 0.5 * ws, 0.5 * hs], dim=-1)


def grid_anchors(featmap_size, stride: int, base: torch.Tensor) -> torch.Tensor:
    """single_level_grid_priors (anchor_generator.py:259-301): (H*W*A, 4), x fastest then anchors."""
    fh, fw = featmap_size
    sx = torch.arange(0, fw, dtype=torch.float32) * stride
    sy = torch.arange(0, fh, dtype=torch.float32) * stride
    xx = sx.repeat(fh)
    yy = sy.view(-1, 1).repeat(1, fw).view(-1)
    shifts = torch.stack([xx, yy, xx, yy], dim=-1)
    return (base[None, :, :] + shifts[:, None, :]).view(-1, 4)


def delta2bbox(rois: torch.Tensor, deltas: torch.Tensor, stds, max_shape, wh_ratio_clip: float = 16 / 1000):
    """delta_xywh_bbox_coder.py:325-359 (means 0)."""
    nb, ncls = deltas.size(0), deltas.size(1) // 4
    if nb == 0:
        return deltas
    d = deltas.reshape(-1, 4) * deltas.new_tensor(stds).view(1, -1)
    r = rois.repeat(1, ncls).reshape(-1, 4)
    pxy = (r[:, :2] + r[:, 2:]) * 0.5
    pwh = r[:, 2:] - r[:, :2]
    dxy_wh = pwh * d[:, :2]
    mr = np.abs(np.log(wh_ratio_clip))
    dwh = d[:, 2:].clamp(min=-mr, max=mr)
    gxy = pxy + dxy_wh
    gwh = pwh * dwh.exp()
    b = torch.cat([gxy - gwh * 0.5, gxy + gwh * 0.5], dim=-1)
    if max_shape is not None:
        b[..., 0::2].clamp_(min=0, max=max_shape[1])
        b[..., 1::2].clamp_(min=0, max=max_shape[0])
    return b.reshape(nb, -1)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_thr: float,
                split_thr: int = 10000):
    """mmcv.ops.batched_nms semantics: offset boxes by idx * (max + 1), one NMS; per-class loop
    at >= split_thr boxes.  Returns (dets[k,5], keep) ordered by descending score."""
    if boxes.numel() == 0:
        return torch.cat([boxes, scores[:, None]], -1), boxes.new_zeros(0, dtype=torch.long)
    mx = boxes.max()
    off = idxs.to(boxes) * (mx + boxes.new_tensor(1))
    bfn = boxes + off[:, None]
    if bfn.shape[0] < split_thr:
        keep = tvops.nms(bfn, scores, iou_thr)
        b, s = boxes[keep], scores[keep]
    else:
        total = scores.new_zeros(scores.size(), dtype=torch.bool)
        after = scores.new_zeros(scores.size())
        for i in torch.unique(idxs):
            m = (idxs == i).nonzero(as_tuple=False).view(-1)
            k = tvops.nms(bfn[m], scores[m], iou_thr)
            total[m[k]] = True
            after[m[k]] = scores[m][k]
        keep = total.nonzero(as_tuple=False).view(-1)
        s, inds = after[keep].sort(descending=True)
        keep = keep[inds]
        b = boxes[keep]
    return torch.cat([b, s[:, None]], -1), keep


def rpn_forward(sd: dict, feats, prefix: str = "rpn_head."):
    """RPNHead.forward_single per level (rpn_head.py:80-97)."""
    out = []
    for x in feats:
        y = F.relu(F.conv2d(x, sd[prefix + "rpn_conv.weight"], sd[prefix + "rpn_conv.bias"], padding=1))
        out.append((F.conv2d(y, sd[prefix + "rpn_cls.weight"], sd[prefix + "rpn_cls.bias"]),
                    F.conv2d(y, sd[prefix + "rpn_reg.weight"], sd[prefix + "rpn_reg.bias"])))
    return out


def rpn_predict_single(cls_list, reg_list, priors_list, img_shape, nms_pre=1000, max_per_img=1000,
                       iou_thr=0.7, min_bbox_size=0):
    """RPNHead._predict_by_feat_single + _bbox_post_process (rpn_head.py:134-304), one image."""
    preds, priors, scores, lvl = [], [], [], []
    for li, (c, r, pr) in enumerate(zip(cls_list, reg_list, priors_list)):
        r = r.permute(1, 2, 0).reshape(-1, 4)
        s = c.permute(1, 2, 0).reshape(-1, 1).sigmoid().squeeze(-1)
        if 0 < nms_pre < s.shape[0]:
            rs, ri = s.sort(descending=True)
            ti = ri[:nms_pre]
            s, r, pr = rs[:nms_pre], r[ti], pr[ti]
        preds.append(r); priors.append(pr); scores.append(s)
        lvl.append(s.new_full((s.size(0),), li, dtype=torch.long))
    boxes = delta2bbox(torch.cat(priors), torch.cat(preds), (1.0, 1.0, 1.0, 1.0), img_shape)
    scores, lvl = torch.cat(scores), torch.cat(lvl)
    if min_bbox_size >= 0:
        w, h = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
        valid = (w > min_bbox_size) & (h > min_bbox_size)
        if not valid.all():
            boxes, scores, lvl = boxes[valid], scores[valid], lvl[valid]
    if boxes.numel() == 0:
        return boxes.new_zeros(0, 4), scores.new_zeros(0)
    dets, keep = batched_nms(boxes.float(), scores.float(), lvl, iou_thr)
    return boxes[keep][:max_per_img], dets[:, -1][:max_per_img]


# --------------------------------------------------------------------------------------------
# RoI head
# --------------------------------------------------------------------------------------------
def sine_positional_encoding(B: int, H: int, W: int, num_feats: int = 128, temperature: int = 10000,
                             scale: float = 2 * math.pi, eps: float = 1e-6) -> torch.Tensor:
    """SinePositionalEncoding.forward with an all-valid mask, normalize=True (positional_encoding.py:60-110)."""
    y = torch.arange(1, H + 1, dtype=torch.float32).view(1, H, 1).repeat(B, 1, W)
    x = torch.arange(1, W + 1, dtype=torch.float32).view(1, 1, W).repeat(B, H, 1)
    y = y / (y[:, -1:, :] + eps) * scale
    x = x / (x[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    px = x[:, :, :, None] / dim_t
    py = y[:, :, :, None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).view(B, H, W, -1)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).view(B, H, W, -1)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


def add_extra_pe(feats):
    """RSPrompterAnchorRoIPromptHead.predict extra_pe branch (M:1566-1574)."""
    bs, c, h, w = feats[0].shape
    pe = sine_positional_encoding(bs, h, w, c // 2)
    return [f + F.interpolate(pe, size=f.shape[-2:], mode="bilinear", align_corners=False) for f in feats]


def map_roi_levels(rois: torch.Tensor, num_levels: int, finest_scale: int = 56) -> torch.Tensor:
    """SingleRoIExtractor.map_roi_levels (single_level_roi_extractor.py:55-62)."""
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lv = torch.floor(torch.log2(scale / finest_scale + 1e-6))
    return lv.clamp(min=0, max=num_levels - 1).long()


def roi_extract(feats, rois: torch.Tensor, out_size: int, strides=(4, 8, 16, 32)) -> torch.Tensor:
    """SingleRoIExtractor.forward (single_level_roi_extractor.py:75-119)."""
    nl = len(strides)
    out = feats[0].new_zeros(rois.size(0), feats[0].shape[1], out_size, out_size)
    lv = map_roi_levels(rois, nl)
    for i in range(nl):
        idx = (lv == i).nonzero(as_tuple=False).squeeze(1)
        if idx.numel() > 0:
            out[idx] = tvops.roi_align(feats[i], rois[idx], out_size, spatial_scale=1.0 / strides[i],
                                       sampling_ratio=0, aligned=True)
    return out


def bbox_head_forward(sd: dict, roi_feats: torch.Tensor, prefix: str = "roi_head.bbox_head."):
    """Shared2FCBBoxHead.forward (convfc_bbox_head.py:163-225): flatten, 2 x (FC + ReLU), fc_cls / fc_reg."""
    x = roi_feats.flatten(1)
    for i in range(2):
        x = F.relu(F.linear(x, sd[f"{prefix}shared_fcs.{i}.weight"], sd[f"{prefix}shared_fcs.{i}.bias"]))
    return (F.linear(x, sd[prefix + "fc_cls.weight"], sd[prefix + "fc_cls.bias"]),
            F.linear(x, sd[prefix + "fc_reg.weight"], sd[prefix + "fc_reg.bias"]))


def bbox_predict_single(roi: torch.Tensor, cls_score: torch.Tensor, bbox_pred: torch.Tensor, img_shape,
                        num_classes: int, score_thr=0.05, iou_thr=0.5, max_per_img=100):
    """BBoxHead._predict_by_feat_single + multiclass_nms (bbox_head.py:476-571, bbox_nms.py:13-105)."""
    scores = F.softmax(cls_score, dim=-1)
    n = roi.size(0)
    r = roi.repeat_interleave(num_classes, dim=0)
    boxes = delta2bbox(r[:, 1:], bbox_pred.view(-1, 4), (0.1, 0.1, 0.2, 0.2), img_shape).view(n, -1)
    b = boxes.view(n, -1, 4).reshape(-1, 4)
    s = scores[:, :-1].reshape(-1)
    labels = torch.arange(num_classes).view(1, -1).expand(n, num_classes).reshape(-1)
    inds = (s > score_thr).nonzero(as_tuple=False).squeeze(1)
    b, s, labels = b[inds], s[inds], labels[inds]
    if b.numel() == 0:
        return b.new_zeros(0, 4), s.new_zeros(0), labels
    dets, keep = batched_nms(b, s, labels, iou_thr)
    dets, keep = dets[:max_per_img], keep[:max_per_img]
    return dets[:, :4], dets[:, 4], labels[keep]


def mask_head_prompts(sd: dict, mask_feats: torch.Tensor, per_point: int = 5,
                      prefix: str = "roi_head.mask_head.") -> torch.Tensor:
    """RSPrompterAnchorMaskHead point_emb + sin fold (M:1641-1651, 1669-1672) -> (N, P, C)."""
    p = prefix + "point_emb."
    x = F.conv2d(mask_feats, sd[p + "0.weight"], sd[p + "0.bias"], stride=2, padding=1)
    x = F.relu(_bn_eval(sd, p + "1", x)).flatten(1)
    x = F.relu(F.linear(x, sd[p + "4.weight"], sd[p + "4.bias"]))
    x = F.relu(F.linear(x, sd[p + "6.weight"], sd[p + "6.bias"]))
    x = F.linear(x, sd[p + "8.weight"], sd[p + "8.bias"])
    x = x.reshape(x.shape[0], per_point, -1)
    return torch.sin(x[..., ::2]) + x[..., 1::2]


def mask_postprocess(mask_logits: torch.Tensor, size, thr: float = 0.5) -> torch.Tensor:
    """RSPrompterAnchorMaskHead._predict_by_feat_single with scale_factor 1 and ori == batch shape
    (M:1746-1784): sigmoid -> bilinear to the input size -> (identity resize) -> >= thr."""
    m = F.interpolate(mask_logits.sigmoid(), size=size, mode="bilinear", align_corners=False).squeeze(1)
    m = F.interpolate(m.unsqueeze(1), size=size, mode="bilinear", align_corners=False).squeeze(1)
    return m >= thr


def mask_postprocess_rescale(mask_logits: torch.Tensor, bboxes: torch.Tensor, meta: dict, thr: float = 0.5):
    """RSPrompterAnchorMaskHead._predict_by_feat_single with rescale=True (M:1746-1784) for one image:
    boxes back to the original image (/ scale_factor), masks: sigmoid -> bilinear to batch_input_shape ->
    crop to the resized (unpadded) image -> bilinear to ori_shape -> >= thr."""
    sf_w, sf_h = meta["scale_factor"]
    scale = bboxes.new_tensor([sf_w, sf_h]).repeat(1, 2)
    boxes = bboxes / scale
    img_h, img_w = meta["ori_shape"][:2]
    m = F.interpolate(mask_logits.sigmoid(), size=tuple(meta["batch_input_shape"]), mode="bilinear",
                      align_corners=False).squeeze(1)
    m = m[:, :int(img_h * sf_h), :int(img_w * sf_w)]
    m = F.interpolate(m.unsqueeze(1), size=(img_h, img_w), mode="bilinear", align_corners=False).squeeze(1)
    return m >= thr, boxes


# --------------------------------------------------------------------------------------------
# whole detector
# --------------------------------------------------------------------------------------------
def _sub(sd: dict, prefix: str) -> dict:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def anchor_predict(sd: dict, vision_arch, decoder_arch, images: torch.Tensor, num_classes: int,
                   select_layers, strides=(4, 8, 16, 32, 64), scales=(4, 8), ratios=(0.5, 1.0, 2.0),
                   points: int = 5, timings: dict | None = None, pseudo_neck: bool = False,
                   extra_boxes: list | None = None):
    """RSPrompterAnchor.predict (M:148-170) for metainfo img_shape == ori_shape == batch shape,
    scale_factor 1: returns a list of per-image dicts(bboxes, scores, labels, masks, mask_logits).
    pseudo_neck: the *-peft-512 configs (MMPretrainSamVisionEncoder + PseudoFeatureAggregator, M:944-984): the neck
    consumes the image embedding instead of the hidden states.
    extra_boxes (per-image [n, 4] tensors): the mask branch (M:1511-1550, 1659-1698) is evaluated a second time on these
    boxes -> results[b]["extra_mask_logits"]; lets a test compare mask logits conditioned on identical box decisions."""
    import time
    t0 = time.perf_counter()
    B, _, H, W = images.shape
    emb, hidden = restate.vit_encoder(_sub(sd, "backbone.vision_encoder."), vision_arch, images)
    t1 = time.perf_counter()
    if pseudo_neck:
        agg = pseudo_feature_aggregator(_sub(sd, "neck.feature_aggregator."), emb)
    else:
        agg = feature_aggregator(_sub(sd, "neck.feature_aggregator."), hidden, list(select_layers))
    feats = simple_fpn(_sub(sd, "neck.feature_spliter."), agg)
    pe = restate.image_wide_positional_embedding(
        sd["shared_image_embedding.shared_image_embedding.positional_embedding"], emb.shape[-1])
    t2 = time.perf_counter()
    heads = rpn_forward(_sub(sd, "rpn_head."), feats, prefix="")
    priors = [grid_anchors(f.shape[-2:], s, base_anchors(s, scales, ratios)) for f, s in zip(feats, strides)]
    props = []
    for b in range(B):
        pb, _ = rpn_predict_single([c[b] for c, _ in heads], [r[b] for _, r in heads], priors, (H, W))
        props.append(pb)
    t3 = time.perf_counter()
    feats_pe = add_extra_pe(feats)
    rois = torch.cat([torch.cat([pb.new_full((pb.shape[0], 1), b), pb], dim=1) for b, pb in enumerate(props)])
    roi_feats = roi_extract(feats_pe[:4], rois, 7)
    cls, reg = bbox_head_forward(_sub(sd, "roi_head.bbox_head."), roi_feats, prefix="")
    results, off = [], 0
    for b, pb in enumerate(props):
        n = pb.shape[0]
        db, ds, dl = bbox_predict_single(rois[off:off + n], cls[off:off + n], reg[off:off + n], (H, W), num_classes)
        off += n
        results.append(dict(bboxes=db, scores=ds, labels=dl))
    t4 = time.perf_counter()
    mrois = torch.cat([torch.cat([r["bboxes"].new_full((r["bboxes"].shape[0], 1), b), r["bboxes"]], dim=1)
                       for b, r in enumerate(results)])
    if mrois.shape[0] > 0:
        mfeats = roi_extract(feats_pe[:4], mrois, 14)
        msd = _sub(sd, "roi_head.mask_head.")
        sparse = mask_head_prompts(msd, mfeats, points, prefix="")
        ids = mrois[:, 0].long()
        dense = sd["roi_head.mask_head.no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(
            ids.numel(), -1, emb.shape[-2], emb.shape[-1])
        logits, _ = restate.mask_decoder(_sub(msd, "mask_decoder.mask_decoder."), decoder_arch, emb[ids],
                                         pe.expand(ids.numel(), -1, -1, -1), sparse[:, None], dense, False)
        logits = logits[:, 0]
        off = 0
        for r in results:
            n = r["bboxes"].shape[0]
            r["mask_logits"] = logits[off:off + n]
            r["masks"] = mask_postprocess(logits[off:off + n], (H, W))
            off += n
    if extra_boxes is not None:
        xrois = torch.cat([torch.cat([bx.new_full((bx.shape[0], 1), b), bx], dim=1) for b, bx in enumerate(extra_boxes)])
        if xrois.shape[0] > 0:
            msd = _sub(sd, "roi_head.mask_head.")
            xsparse = mask_head_prompts(msd, roi_extract(feats_pe[:4], xrois, 14), points, prefix="")
            xids = xrois[:, 0].long()
            xdense = sd["roi_head.mask_head.no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(
                xids.numel(), -1, emb.shape[-2], emb.shape[-1])
            xl, _ = restate.mask_decoder(_sub(msd, "mask_decoder.mask_decoder."), decoder_arch, emb[xids],
                                         pe.expand(xids.numel(), -1, -1, -1), xsparse[:, None], xdense, False)
            off = 0
            for r, bx in zip(results, extra_boxes):
                r["extra_mask_logits"] = xl[off:off + bx.shape[0], 0]
                off += bx.shape[0]
    t5 = time.perf_counter()
    if timings is not None:
        timings.update(encoder=t1 - t0, neck=t2 - t1, rpn=t3 - t2, roi=t4 - t3, mask=t5 - t4)
    return results


# ------------------------------------------------------------------------------ stock Mask R-CNN mask branch (SAMSegMaskRCNN)
def fcn_mask_head(sd: dict, roi_feats: torch.Tensor, prefix: str = "roi_head.mask_head.") -> torch.Tensor:
    """FCNMaskHead.forward (fcn_mask_head.py:128-147): convs (conv3x3 + ReLU) -> deconv k2 s2 + ReLU -> 1x1 logits.
    roi_feats fp32 [n, C, 14, 14] -> [n, num_classes, 28, 28]."""
    x = roi_feats
    i = 0
    while f"{prefix}convs.{i}.conv.weight" in sd:
        x = F.relu(F.conv2d(x, sd[f"{prefix}convs.{i}.conv.weight"], sd[f"{prefix}convs.{i}.conv.bias"], padding=1))
        i += 1
    x = F.relu(F.conv_transpose2d(x, sd[prefix + "upsample.weight"], sd[prefix + "upsample.bias"], stride=2))
    return F.conv2d(x, sd[prefix + "conv_logits.weight"], sd[prefix + "conv_logits.bias"])


def paste_masks_in_boxes(probs: torch.Tensor, boxes: torch.Tensor, img_h: int, img_w: int) -> torch.Tensor:
    """_do_paste_mask (fcn_mask_head.py:395-458, skip_empty=False) written out: every image pixel centre is mapped into
    the box's [-1, 1] frame and the RoI mask is sampled bilinearly with F.grid_sample's align_corners=False / zero
    padding rule (pixel index = ((g + 1) * size - 1) / 2; taps outside the grid contribute 0).
    probs fp32 [n, hm, wm], boxes fp32 [n, 4] -> fp32 [n, img_h, img_w]."""
    n, hm, wm = probs.shape
    x0, y0, x1, y1 = boxes[:, 0:1], boxes[:, 1:2], boxes[:, 2:3], boxes[:, 3:4]
    ys = torch.arange(img_h, dtype=torch.float32) + 0.5
    xs = torch.arange(img_w, dtype=torch.float32) + 0.5
    gy = (ys[None] - y0) / (y1 - y0) * 2 - 1          # [n, H]
    gx = (xs[None] - x0) / (x1 - x0) * 2 - 1          # [n, W]
    gy = torch.where(torch.isinf(gy), torch.zeros_like(gy), gy)
    gx = torch.where(torch.isinf(gx), torch.zeros_like(gx), gx)
    iy = ((gy + 1) * hm - 1) / 2
    ix = ((gx + 1) * wm - 1) / 2
    fy, fx = torch.floor(iy), torch.floor(ix)
    out = torch.zeros(n, img_h, img_w)
    ar = torch.arange(n)[:, None, None]
    for dy in (0, 1):
        yy = fy + dy
        wy = ((fy + 1) - iy) if dy == 0 else (iy - fy)
        vy = (yy >= 0) & (yy < hm)
        for dx in (0, 1):
            xx = fx + dx
            wx = ((fx + 1) - ix) if dx == 0 else (ix - fx)
            vx = (xx >= 0) & (xx < wm)
            tap = probs[ar, yy.clamp(0, hm - 1).long()[:, :, None], xx.clamp(0, wm - 1).long()[:, None, :]]
            out += tap * (wy[:, :, None] * wx[:, None, :]) * (vy[:, :, None] & vx[:, None, :])
    return out


def fcn_mask_predict_single(mask_logits: torch.Tensor, bboxes: torch.Tensor, labels: torch.Tensor, ori_hw,
                            scale_factor=(1.0, 1.0), rescale: bool = True, thr: float = 0.5):
    """FCNMaskHead._predict_by_feat_single (fcn_mask_head.py:278-393): sigmoid, boxes to the output frame, the label's
    channel, paste, >= thr.  -> (bool masks [n, h, w], boxes in the output frame)."""
    probs = torch.sigmoid(mask_logits)
    sf = bboxes.new_tensor(scale_factor).repeat(2)
    img_h, img_w = int(ori_hw[0]), int(ori_hw[1])
    if rescale:
        bboxes = bboxes / sf
    else:
        img_h, img_w = int(round(img_h * float(sf[1]))), int(round(img_w * float(sf[0])))
    n = probs.shape[0]
    if n == 0:
        return torch.zeros(0, img_h, img_w, dtype=torch.bool), bboxes
    sel = probs[torch.arange(n), labels]
    return paste_masks_in_boxes(sel, bboxes, img_h, img_w) >= thr, bboxes


def maskrcnn_predict(sd: dict, vision_arch, images: torch.Tensor, num_classes: int, select_layers,
                     strides=(4, 8, 16, 32, 64), scales=(8,), ratios=(0.5, 1.0, 2.0), extra_boxes: list | None = None):
    """SAMSegMaskRCNN.predict (M:1218-1244 + two_stage.py:196-243) for img_shape == ori_shape == batch shape, scale 1:
    per-image dicts(bboxes, scores, labels, masks, mask_logits [n, num_classes, 28, 28]).
    extra_boxes / extra_labels: see anchor_predict - the mask branch evaluated on given detections."""
    B, _, H, W = images.shape
    emb, hidden = restate.vit_encoder(_sub(sd, "backbone.vision_encoder."), vision_arch, images)
    agg = feature_aggregator(_sub(sd, "neck.feature_aggregator."), hidden, list(select_layers))
    feats = simple_fpn(_sub(sd, "neck.feature_spliter."), agg)
    heads = rpn_forward(_sub(sd, "rpn_head."), feats, prefix="")
    priors = [grid_anchors(f.shape[-2:], s, base_anchors(s, scales, ratios)) for f, s in zip(feats, strides)]
    props = [rpn_predict_single([c[b] for c, _ in heads], [r[b] for _, r in heads], priors, (H, W))[0] for b in range(B)]
    rois = torch.cat([torch.cat([pb.new_full((pb.shape[0], 1), b), pb], dim=1) for b, pb in enumerate(props)])
    cls, reg = bbox_head_forward(_sub(sd, "roi_head.bbox_head."), roi_extract(feats[:4], rois, 7), prefix="")
    results, off = [], 0
    msd = _sub(sd, "roi_head.mask_head.")
    for b, pb in enumerate(props):
        n = pb.shape[0]
        db, ds, dl = bbox_predict_single(rois[off:off + n], cls[off:off + n], reg[off:off + n], (H, W), num_classes)
        off += n
        r = dict(bboxes=db, scores=ds, labels=dl, proposals=pb)
        if db.shape[0] > 0:
            mrois = torch.cat([db.new_full((db.shape[0], 1), b), db], dim=1)
            r["mask_logits"] = fcn_mask_head(msd, roi_extract(feats[:4], mrois, 14), prefix="")
            r["masks"], _ = fcn_mask_predict_single(r["mask_logits"], db, dl, (H, W))
        if extra_boxes is not None and extra_boxes[b].shape[0] > 0:
            xb = extra_boxes[b]
            xr = torch.cat([xb.new_full((xb.shape[0], 1), b), xb], dim=1)
            r["extra_mask_logits"] = fcn_mask_head(msd, roi_extract(feats[:4], xr, 14), prefix="")
        results.append(r)
    return results


def mask2bbox(masks: torch.Tensor) -> torch.Tensor:
    """Tight boxes of boolean masks (mmdet/structures/mask/utils.py:56-77)."""
    n = masks.shape[0]
    out = masks.new_zeros((n, 4), dtype=torch.float32)
    xa, ya = torch.any(masks, dim=1), torch.any(masks, dim=2)
    for i in range(n):
        x, y = torch.where(xa[i])[0], torch.where(ya[i])[0]
        if len(x) > 0 and len(y) > 0:
            out[i] = out.new_tensor([x[0], y[0], x[-1] + 1, y[-1] + 1])
    return out
