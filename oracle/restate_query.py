"""fp32 CPU restatement of the RSPrompter-query head path (TEST INFRASTRUCTURE, see oracle/__init__.py):
MSDeformAttnPixelDecoder -> Mask2Former transformer decoder -> RSMask2FormerHead._forward_head (SAM
decoder with per-prompt dense embeddings) -> RSMaskFormerFusionHead.instance_postprocess.

References: mmdet/models/layers/msdeformattn_pixel_decoder.py:144-246, layers/transformer/
{mask2former_layers.py:9-135, detr_layers.py:84-374, deformable_detr_layers.py:21-119,237-249},
mmdet/rsprompter/models.py (M:) 274-463, 633-715, seg_heads/panoptic_fusion_heads/
maskformer_fusion_head.py:126-182.  mmcv bricks absent from /root/reference are restated from their
documented semantics: MultiScaleDeformableAttention (pure-PyTorch grid_sample form),
mmcv MultiheadAttention wrapper (adds query_pos / key_pos, returns identity + out), FFN (identity +
out), ConvModule(norm=GN) ordering conv -> norm -> act.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import restate
from .restate_anchor import mask2bbox, sine_positional_encoding


def _conv_gn(sd: dict, p: str, x: torch.Tensor, padding: int = 0, relu: bool = False, groups: int = 32):
    """mmcv ConvModule(norm_cfg=GN): conv (bias only when given) -> GroupNorm -> optional ReLU."""
    x = F.conv2d(x, sd[p + "conv.weight"], sd.get(p + "conv.bias"), padding=padding)
    x = F.group_norm(x, groups, sd[p + "gn.weight"], sd[p + "gn.bias"], 1e-5)
    return F.relu(x) if relu else x


def ms_deform_attn(sd: dict, p: str, query, query_pos, reference_points, spatial_shapes, heads=8, points=4):
    """mmcv MultiScaleDeformableAttention.forward (batch_first), value = identity = query."""
    B, nq, E = query.shape
    L = len(spatial_shapes)
    identity = query
    q = query + query_pos
    value = F.linear(query, sd[p + "value_proj.weight"], sd[p + "value_proj.bias"]).view(B, nq, heads, E // heads)
    off = F.linear(q, sd[p + "sampling_offsets.weight"], sd[p + "sampling_offsets.bias"]).view(B, nq, heads, L, points, 2)
    aw = F.linear(q, sd[p + "attention_weights.weight"], sd[p + "attention_weights.bias"]).view(B, nq, heads, L * points)
    aw = aw.softmax(-1).view(B, nq, heads, L, points)
    norm = torch.tensor([[w, h] for h, w in spatial_shapes], dtype=query.dtype)
    loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    grids = 2 * loc - 1
    vals = value.split([h * w for h, w in spatial_shapes], dim=1)
    sampled = []
    for lvl, (h, w) in enumerate(spatial_shapes):
        v = vals[lvl].flatten(2).transpose(1, 2).reshape(B * heads, E // heads, h, w)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    aw = aw.transpose(1, 2).reshape(B * heads, 1, nq, L * points)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).view(B, E, nq).transpose(1, 2)
    out = F.linear(out, sd[p + "output_proj.weight"], sd[p + "output_proj.bias"])
    return out + identity


def _ffn(sd: dict, p: str, x):
    """mmcv FFN(num_fcs=2, ReLU): identity + Linear(ReLU(Linear(x)))."""
    y = F.relu(F.linear(x, sd[p + "layers.0.0.weight"], sd[p + "layers.0.0.bias"]))
    return x + F.linear(y, sd[p + "layers.1.weight"], sd[p + "layers.1.bias"])


def _ln(sd: dict, p: str, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], eps)


def pixel_decoder(sd: dict, feats, prefix="", enc_levels=3, enc_layers=3, heads=8, points=4, num_outs=3):
    """MSDeformAttnPixelDecoder.forward (msdeformattn_pixel_decoder.py:144-246).
    feats: 5 NCHW maps (high -> low resolution).  -> (mask_feature, [memories low -> high resolution])."""
    p = prefix
    nl = len(feats)
    B = feats[0].shape[0]
    E = sd[p + "level_encoding.weight"].shape[1]
    inputs, poss, shapes, refs = [], [], [], []
    for i in range(enc_levels):
        f = feats[nl - 1 - i]
        h, w = f.shape[-2:]
        proj = _conv_gn(sd, f"{p}input_convs.{i}.", f)
        pos = sine_positional_encoding(B, h, w, E // 2) + sd[p + "level_encoding.weight"][i].view(1, -1, 1, 1)
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        ref = torch.stack([(xs.reshape(-1) + 0.5) / w, (ys.reshape(-1) + 0.5) / h], dim=-1)
        inputs.append(proj.flatten(2).permute(0, 2, 1)); poss.append(pos.flatten(2).permute(0, 2, 1))
        shapes.append((h, w)); refs.append(ref)
    x = torch.cat(inputs, dim=1)
    pos = torch.cat(poss, dim=1)
    ref = torch.cat(refs, dim=0)[None, :, None].repeat(B, 1, enc_levels, 1)
    for l in range(enc_layers):
        lp = f"{p}encoder.layers.{l}."
        x = ms_deform_attn(sd, lp + "self_attn.", x, pos, ref, shapes, heads, points)
        x = _ln(sd, lp + "norms.0.", x)
        x = _ffn(sd, lp + "ffn.", x)
        x = _ln(sd, lp + "norms.1.", x)
    mem = x.permute(0, 2, 1)
    outs = [m.reshape(B, E, h, w) for m, (h, w) in zip(torch.split(mem, [h * w for h, w in shapes], dim=-1), shapes)]
    for i in range(nl - enc_levels - 1, -1, -1):
        cur = _conv_gn(sd, f"{p}lateral_convs.{i}.", feats[i])
        y = cur + F.interpolate(outs[-1], size=cur.shape[-2:], mode="bilinear", align_corners=False)
        outs.append(_conv_gn(sd, f"{p}output_convs.{i}.", y, padding=1, relu=True))
    mask_feature = F.conv2d(outs[-1], sd[p + "mask_feature.weight"], sd[p + "mask_feature.bias"])
    return mask_feature, outs[:num_outs]


def _mha(sd: dict, p: str, query, key, value, query_pos, key_pos, attn_mask, heads=8):
    """mmcv MultiheadAttention wrapper around nn.MultiheadAttention (batch_first): identity + attn."""
    identity = query
    q = query + query_pos if query_pos is not None else query
    k = key + key_pos if key_pos is not None else key
    B, nq, E = q.shape
    W, b = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
    qq = F.linear(q, W[:E], b[:E]).view(B, nq, heads, -1).transpose(1, 2)
    kk = F.linear(k, W[E:2 * E], b[E:2 * E]).view(B, k.shape[1], heads, -1).transpose(1, 2)
    vv = F.linear(value, W[2 * E:], b[2 * E:]).view(B, k.shape[1], heads, -1).transpose(1, 2)
    s = (qq @ kk.transpose(-2, -1)) * (qq.shape[-1] ** -0.5)
    if attn_mask is not None:
        s = s.masked_fill(attn_mask.view(B, heads, nq, -1), float("-inf"))
    o = (s.softmax(-1) @ vv).transpose(1, 2).reshape(B, nq, E)
    return identity + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])


def _mlp(sd: dict, p: str, idx, x):
    for j, i in enumerate(idx):
        x = F.linear(x, sd[f"{p}{i}.weight"], sd[f"{p}{i}.bias"])
        if j < len(idx) - 1:
            x = F.relu(x)
    return x


def forward_head(sd: dict, darch, dec_sd: dict, pe_sd: dict, decoder_out, mask_feature, target_size, emb, pos,
                 points=5, heads=8, run_sam=True, prefix=""):
    """RSMask2FormerHead._forward_head with decoder_plus=True (M:332-393)."""
    p = prefix
    B, nq, _ = decoder_out.shape
    x = _ln(sd, p + "transformer_decoder.post_norm.", decoder_out)
    cls_pred = _mlp(sd, p + "cls_embed.", [0, 2], x)
    mask_embed = _mlp(sd, p + "mask_embed.", [0, 2, 4], x)
    mask_pred_plus = torch.einsum("bqc,bchw->bqhw", mask_embed, mask_feature)
    mask_pred = None
    if run_sam:
        pts = _mlp(sd, p + "point_emb.", [0, 2, 4], x).reshape(B, nq, points, -1)
        pts = torch.sin(pts[..., ::2]) + pts[..., 1::2]
        sparse = pts.reshape(B * nq, 1, points, -1)
        dense = restate.sam_mask_embedding(pe_sd, mask_pred_plus.reshape(B * nq, 1, *mask_pred_plus.shape[-2:]))
        e = torch.repeat_interleave(emb, nq, dim=0)
        pp = torch.repeat_interleave(pos, nq, dim=0)
        m, _ = restate.mask_decoder(dec_sd, darch, e, pp, sparse, dense, False)
        mask_pred = m.reshape(B, nq, *m.shape[-2:])
    am = F.interpolate(mask_pred_plus, target_size, mode="bilinear", align_corners=False)
    am = am.flatten(2).unsqueeze(1).repeat(1, heads, 1, 1).flatten(0, 1)
    am = am.sigmoid() < 0.5
    return cls_pred, mask_pred, am, mask_pred_plus


def mask2former_head(sd: dict, darch, dec_sd, pe_sd, feats, emb, pos, num_layers=6, levels=3, heads=8, points=5,
                     prefix="", sam_every_layer=False):
    """RSMask2FormerHead.forward (M:395-463) -> last-layer (cls, mask_pred, mask_pred_plus).
    The SAM decoder of the intermediate layers is dead code at inference with decoder_plus=True
    (attn masks come from mask_pred_plus, predict() reads only the last entry, M:380-385,644-645)."""
    p = prefix
    B = feats[0].shape[0]
    mask_feature, memories = pixel_decoder(sd, feats, p + "pixel_decoder.")
    E = memories[0].shape[1]
    dec_in, dec_pos = [], []
    for i in range(levels):
        m = memories[i]
        dec_in.append(m.flatten(2).permute(0, 2, 1) + sd[p + "level_embed.weight"][i].view(1, 1, -1))
        dec_pos.append(sine_positional_encoding(B, m.shape[-2], m.shape[-1], E // 2).flatten(2).permute(0, 2, 1))
    qf = sd[p + "query_feat.weight"].unsqueeze(0).repeat(B, 1, 1)
    qe = sd[p + "query_embed.weight"].unsqueeze(0).repeat(B, 1, 1)
    cls, mp, am, mpp = forward_head(sd, darch, dec_sd, pe_sd, qf, mask_feature, memories[0].shape[-2:], emb, pos,
                                    points, heads, run_sam=sam_every_layer or num_layers == 0, prefix=p)
    for i in range(num_layers):
        lvl = i % levels
        am = am & (am.sum(-1) != am.shape[-1]).unsqueeze(-1)
        lp = f"{p}transformer_decoder.layers.{i}."
        qf = _mha(sd, lp + "cross_attn.", qf, dec_in[lvl], dec_in[lvl], qe, dec_pos[lvl], am, heads)
        qf = _ln(sd, lp + "norms.0.", qf)
        qf = _mha(sd, lp + "self_attn.", qf, qf, qf, qe, qe, None, heads)
        qf = _ln(sd, lp + "norms.1.", qf)
        qf = _ffn(sd, lp + "ffn.", qf)
        qf = _ln(sd, lp + "norms.2.", qf)
        last = i == num_layers - 1
        cls, mp, am, mpp = forward_head(sd, darch, dec_sd, pe_sd, qf, mask_feature,
                                        memories[(i + 1) % levels].shape[-2:], emb, pos, points, heads,
                                        run_sam=sam_every_layer or last, prefix=p)
    return cls, mp, mpp


def instance_postprocess(mask_cls: torch.Tensor, mask_pred: torch.Tensor, num_classes: int, max_per_image=100):
    """MaskFormerFusionHead.instance_postprocess (maskformer_fusion_head.py:126-182), one image;
    mask_pred already at output resolution."""
    nq = mask_cls.shape[0]
    scores = F.softmax(mask_cls, dim=-1)[:, :-1]
    labels = torch.arange(num_classes).unsqueeze(0).repeat(nq, 1).flatten(0, 1)
    sc, top = scores.flatten(0, 1).topk(max_per_image, sorted=False)
    lab = labels[top]
    qidx = top // num_classes
    mp = mask_pred[qidx]
    binary = (mp > 0).float()
    mscore = (mp.sigmoid() * binary).flatten(1).sum(1) / (binary.flatten(1).sum(1) + 1e-6)
    return dict(bboxes=mask2bbox(binary.bool()), labels=lab, scores=sc * mscore, masks=binary.bool(), query=qidx)


def fusion_rescale(mask_pred: torch.Tensor, meta: dict, rescale: bool = True) -> torch.Tensor:
    """RSMask2FormerHead.predict up-sampling + RSMaskFormerFusionHead.predict crop / rescale for one image
    (M:652-656, 679-691): logits [nq, hm, wm] -> batch_input_shape -> crop to the resized image -> ori_shape."""
    m = F.interpolate(mask_pred[None], size=tuple(meta["batch_input_shape"]), mode="bilinear", align_corners=False)[0]
    ori_h, ori_w = meta["ori_shape"][:2]
    sf_w, sf_h = meta["scale_factor"]
    m = m[:, :int(ori_h * sf_h), :int(ori_w * sf_w)]
    if rescale:
        m = F.interpolate(m[:, None], size=(ori_h, ori_w), mode="bilinear", align_corners=False)[:, 0]
    return m


def _sub(sd: dict, prefix: str) -> dict:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def query_predict(sd: dict, vision_arch, decoder_arch, images: torch.Tensor, num_classes: int, select_layers,
                  points: int = 5, max_per_image: int = 100, timings: dict | None = None):
    """RSPrompterQuery.predict (M:249-272) for metainfo img_shape == ori_shape == batch shape, scale_factor 1:
    extract_feat (M:217-234) -> RSMask2FormerHead.predict (M:633-658: last-layer cls / SAM-decoder masks, bilinear to
    the batch shape) -> RSMaskFormerFusionHead.predict (M:663-715, instance_on only) -> per-image dicts(bboxes, scores,
    labels, masks, query, mask_logits [low-res logits of every query], cls)."""
    import time
    from . import restate, restate_anchor
    t0 = time.perf_counter()
    B, _, H, W = images.shape
    emb, hidden = restate.vit_encoder(_sub(sd, "backbone.vision_encoder."), vision_arch, images)
    t1 = time.perf_counter()
    agg = restate_anchor.feature_aggregator(_sub(sd, "neck.feature_aggregator."), hidden, list(select_layers))
    feats = restate_anchor.simple_fpn(_sub(sd, "neck.feature_spliter."), agg)
    pe = restate.image_wide_positional_embedding(
        sd["shared_image_embedding.shared_image_embedding.positional_embedding"], emb.shape[-1])
    t2 = time.perf_counter()
    hsd = _sub(sd, "panoptic_head.")
    dec_sd = _sub(hsd, "mask_decoder.mask_decoder.")
    pe_sd = {k[len("sam_"):]: v for k, v in hsd.items() if k.startswith("sam_mask_embed.")}
    cls, mp, mpp = mask2former_head(hsd, decoder_arch, dec_sd, pe_sd, feats, emb, pe.expand(B, -1, -1, -1), points=points)
    t3 = time.perf_counter()
    up = F.interpolate(mp, size=(H, W), mode="bilinear", align_corners=False)
    out = []
    for b in range(B):
        r = instance_postprocess(cls[b], up[b], num_classes, max_per_image)
        r.update(mask_logits=mp[b], cls=cls[b])
        out.append(r)
    t4 = time.perf_counter()
    if timings is not None:
        timings.update(encoder=t1 - t0, neck=t2 - t1, head=t3 - t2, post=t4 - t3)
    return out


# ------------------------------------------------------------------------------ stock Mask2FormerHead (SAMSegMask2Former)
def stock_forward_head(sd: dict, decoder_out, mask_feature, target_size, heads=8, prefix=""):
    """Mask2FormerHead._forward_head (dense_heads/mask2former_head.py:340-380)."""
    p = prefix
    x = _ln(sd, p + "transformer_decoder.post_norm.", decoder_out)
    cls_pred = F.linear(x, sd[p + "cls_embed.weight"], sd[p + "cls_embed.bias"])
    mask_embed = _mlp(sd, p + "mask_embed.", [0, 2, 4], x)
    mask_pred = torch.einsum("bqc,bchw->bqhw", mask_embed, mask_feature)
    am = F.interpolate(mask_pred, target_size, mode="bilinear", align_corners=False)
    am = am.flatten(2).unsqueeze(1).repeat(1, heads, 1, 1).flatten(0, 1)
    return cls_pred, mask_pred, am.sigmoid() < 0.5


def stock_mask2former_decoder(sd: dict, mask_feature, memories, levels=3, heads=8, prefix=""):
    """The part of Mask2FormerHead.forward behind the pixel decoder (mask2former_head.py:404-460): level / positional
    embeddings, the masked-attention decoder layers, _forward_head after each -> the last layer's (cls_pred, mask_pred)."""
    p = prefix
    B = mask_feature.shape[0]
    E = memories[0].shape[1]
    dec_in, dec_pos = [], []
    for i in range(levels):
        m = memories[i]
        dec_in.append(m.flatten(2).permute(0, 2, 1) + sd[p + "level_embed.weight"][i].view(1, 1, -1))
        dec_pos.append(sine_positional_encoding(B, m.shape[-2], m.shape[-1], E // 2).flatten(2).permute(0, 2, 1))
    qf = sd[p + "query_feat.weight"].unsqueeze(0).repeat(B, 1, 1)
    qe = sd[p + "query_embed.weight"].unsqueeze(0).repeat(B, 1, 1)
    cls, mp, am = stock_forward_head(sd, qf, mask_feature, memories[0].shape[-2:], heads, p)
    i = 0
    while f"{p}transformer_decoder.layers.{i}.norms.0.weight" in sd:
        lvl = i % levels
        am = am & (am.sum(-1) != am.shape[-1]).unsqueeze(-1)
        lp = f"{p}transformer_decoder.layers.{i}."
        qf = _mha(sd, lp + "cross_attn.", qf, dec_in[lvl], dec_in[lvl], qe, dec_pos[lvl], am, heads)
        qf = _ln(sd, lp + "norms.0.", qf)
        qf = _mha(sd, lp + "self_attn.", qf, qf, qf, qe, qe, None, heads)
        qf = _ln(sd, lp + "norms.1.", qf)
        qf = _ffn(sd, lp + "ffn.", qf)
        qf = _ln(sd, lp + "norms.2.", qf)
        cls, mp, am = stock_forward_head(sd, qf, mask_feature, memories[(i + 1) % levels].shape[-2:], heads, p)
        i += 1
    return cls, mp


def stock_mask2former_head(sd: dict, feats, levels=3, heads=8, prefix=""):
    """Mask2FormerHead.forward (mask2former_head.py:382-460) -> the last layer's (cls_pred, mask_pred)."""
    mask_feature, memories = pixel_decoder(sd, feats, prefix + "pixel_decoder.", heads=heads)
    return stock_mask2former_decoder(sd, mask_feature, memories, levels, heads, prefix)


def samseg_mask2former_predict(sd: dict, vision_arch, images: torch.Tensor, num_classes: int, select_layers,
                               max_per_image: int = 100):
    """SAMSegMask2Former.predict (M:1247-1274 + detectors/maskformer.py:95-140) for img_shape == ori_shape == batch
    shape, scale 1: extract_feat -> Mask2FormerHead.predict (maskformer_head.py:569-604: last layer, bilinear to the
    batch shape) -> MaskFormerFusionHead.predict (instance_on) -> per-image dicts as query_predict returns."""
    from . import restate, restate_anchor
    B, _, H, W = images.shape
    emb, hidden = restate.vit_encoder(_sub(sd, "backbone.vision_encoder."), vision_arch, images)
    agg = restate_anchor.feature_aggregator(_sub(sd, "neck.feature_aggregator."), hidden, list(select_layers))
    feats = restate_anchor.simple_fpn(_sub(sd, "neck.feature_spliter."), agg)
    cls, mp = stock_mask2former_head(_sub(sd, "panoptic_head."), feats)
    up = F.interpolate(mp, size=(H, W), mode="bilinear", align_corners=False)
    out = []
    for b in range(B):
        r = instance_postprocess(cls[b], up[b], num_classes, max_per_image)
        r.update(mask_logits=mp[b], cls=cls[b])
        out.append(r)
    return out
