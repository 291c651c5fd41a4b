"""Generates tests/golden/reference_functions.pt by EXECUTING the reference's own in-tree source
(read from /root/reference at generation time; nothing is copied into this repo).

The reference package cannot be imported here (mmcv / mmengine / peft are absent), so the pure
tensor functions / methods on the hot path are pulled out of their files with `ast` and executed
in a namespace that only provides torch / numpy / math; methods get a SimpleNamespace `self`.
Run in the build container:   python tests/golden/make_golden.py
The fixtures pin oracle/restate*.py in tests/test_golden_cpu.py (no /root/reference at test time).
"""
import ast
import math
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_functions.pt")


def load_defs(relpath, names, cls=None):
    """Compile the named top-level functions (or methods of class `cls`) of a reference file."""
    src = open(os.path.join(REF, relpath)).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    ns = dict(torch=torch, F=F, np=np, math=math, Tensor=torch.Tensor, Tuple=tuple, Optional=None,
              Sequence=None, Union=None, List=list)
    for node in body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []
            node.returns = None
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(ast.fix_missing_locations(mod), relpath, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, f"{relpath}: {missing} not found"
    return ns


def main():
    g = torch.Generator().manual_seed(1234)
    fx = {}

    # ---- mmpretrain/models/backbones/vit_sam.py:17-157
    vs = load_defs("mmpretrain/models/backbones/vit_sam.py",
                   ["window_partition", "window_unpartition", "get_rel_pos", "add_decomposed_rel_pos"])
    x = torch.randn(2, 20, 23, 8, generator=g)
    win, pad_hw = vs["window_partition"](x, 14)
    fx["window_partition"] = dict(x=x, ws=14, windows=win, pad_hw=tuple(pad_hw),
                                  back=vs["window_unpartition"](win, 14, pad_hw, (20, 23)))
    table = torch.randn(27, 16, generator=g)
    fx["get_rel_pos_same"] = dict(q=14, k=14, table=table, out=vs["get_rel_pos"](14, 14, table))
    table2 = torch.randn(127, 16, generator=g)
    fx["get_rel_pos_resized"] = dict(q=32, k=32, table=table2, out=vs["get_rel_pos"](32, 32, table2))
    S, hd, nb = 6, 16, 3
    q = torch.randn(nb, S * S, hd, generator=g)
    rh, rw = torch.randn(2 * S - 1, hd, generator=g), torch.randn(2 * S - 1, hd, generator=g)
    attn = torch.zeros(nb, S * S, S * S)
    fx["decomposed_rel_pos"] = dict(q=q, rel_h=rh, rel_w=rw, S=S,
                                    bias=vs["add_decomposed_rel_pos"](attn, q, rh, rw, (S, S), (S, S)))

    # ---- mmdet/models/task_modules/coders/delta_xywh_bbox_coder.py:262-359
    dc = load_defs("mmdet/models/task_modules/coders/delta_xywh_bbox_coder.py", ["delta2bbox"])
    rois = torch.rand(40, 4, generator=g) * 200
    rois[:, 2:] = rois[:, :2] + torch.rand(40, 2, generator=g) * 300 + 1
    deltas = torch.randn(40, 12, generator=g) * 2
    fx["delta2bbox"] = dict(rois=rois, deltas=deltas, stds=(0.1, 0.1, 0.2, 0.2), max_shape=(256, 320),
                            out=dc["delta2bbox"](rois, deltas, (0., 0., 0., 0.), (0.1, 0.1, 0.2, 0.2), (256, 320)))
    d1 = torch.randn(40, 4, generator=g)
    fx["delta2bbox_rpn"] = dict(rois=rois, deltas=d1, stds=(1., 1., 1., 1.), max_shape=(256, 320),
                                out=dc["delta2bbox"](rois, d1, (0., 0., 0., 0.), (1., 1., 1., 1.), (256, 320)))

    # ---- mmdet/models/task_modules/prior_generators/anchor_generator.py:161-301
    ag = load_defs("mmdet/models/task_modules/prior_generators/anchor_generator.py",
                   ["gen_single_level_base_anchors", "_meshgrid", "single_level_grid_priors"], cls="AnchorGenerator")
    self_ = SimpleNamespace(center_offset=0.0, scale_major=True, strides=[(8, 8)], use_box_type=False)
    self_._meshgrid = lambda x, y, row_major=True: ag["_meshgrid"](self_, x, y, row_major)
    base = ag["gen_single_level_base_anchors"](self_, 8, torch.tensor([4., 8.]), torch.tensor([0.5, 1.0, 2.0]))
    self_.base_anchors = [base]
    fx["anchors"] = dict(base_size=8, scales=[4, 8], ratios=[0.5, 1.0, 2.0], base=base, featmap=(3, 5), stride=8,
                         grid=ag["single_level_grid_priors"](self_, (3, 5), 0, torch.float32, "cpu"))

    # ---- mmdet/models/layers/positional_encoding.py:60-110
    pe = load_defs("mmdet/models/layers/positional_encoding.py", ["forward"], cls="SinePositionalEncoding")
    self_ = SimpleNamespace(num_feats=8, temperature=10000, normalize=True, scale=2 * math.pi, eps=1e-6, offset=0.0)
    fx["sine_pe"] = dict(B=2, H=5, W=7, num_feats=8, out=pe["forward"](self_, torch.zeros(2, 5, 7, dtype=torch.bool)))

    # ---- mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:40-62
    re_ = load_defs("mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py", ["map_roi_levels"],
                    cls="SingleRoIExtractor")
    r5 = torch.cat([torch.zeros(40, 1), rois * torch.rand(40, 1, generator=g) * 4], dim=1)
    fx["map_roi_levels"] = dict(rois=r5, num_levels=4,
                                out=re_["map_roi_levels"](SimpleNamespace(finest_scale=56), r5, 4))

    # ---- mmdet/structures/mask/utils.py:56-77
    mu = load_defs("mmdet/structures/mask/utils.py", ["mask2bbox"])
    masks = torch.rand(6, 20, 30, generator=g) > 0.9
    masks[2] = False
    fx["mask2bbox"] = dict(masks=masks, out=mu["mask2bbox"](masks))

    # ---- mmdet/rsprompter/models.py:45-50 LN2d.forward
    ln = load_defs("mmdet/rsprompter/models.py", ["forward"], cls="LN2d")
    w, b = torch.randn(8, generator=g), torch.randn(8, generator=g)
    xx = torch.randn(2, 8, 4, 5, generator=g)
    fx["ln2d"] = dict(x=xx, weight=w, bias=b, eps=1e-6,
                      out=ln["forward"](SimpleNamespace(weight=w, bias=b, eps=1e-6), xx))

    # ---- mmdet/models/seg_heads/panoptic_fusion_heads/maskformer_fusion_head.py:126-182 instance_postprocess
    class _Inst(SimpleNamespace):
        pass
    fh = load_defs("mmdet/models/seg_heads/panoptic_fusion_heads/maskformer_fusion_head.py", ["instance_postprocess"],
                   cls="MaskFormerFusionHead")
    fh["instance_postprocess"].__globals__.update(mask2bbox=mu["mask2bbox"], InstanceData=_Inst)
    nq, ncls = 12, 5
    mask_cls = torch.randn(nq, ncls + 1, generator=g) * 2
    yy, xx2 = torch.meshgrid(torch.arange(40), torch.arange(48), indexing="ij")
    cy, cx, rr = torch.rand(nq, 1, 1, generator=g) * 40, torch.rand(nq, 1, 1, generator=g) * 48, torch.rand(nq, 1, 1, generator=g) * 12
    mask_pred = (rr + 2 - ((yy - cy) ** 2 + (xx2 - cx) ** 2).sqrt()) * 1.3 + 0.1 * torch.randn(nq, 40, 48, generator=g)
    self_ = SimpleNamespace(test_cfg=dict(max_per_image=7), num_classes=ncls, num_things_classes=ncls)
    res = fh["instance_postprocess"](self_, mask_cls, mask_pred)
    fx["instance_postprocess"] = dict(mask_cls=mask_cls, mask_pred=mask_pred, num_classes=ncls, max_per_image=7,
                                      bboxes=res.bboxes, labels=res.labels, scores=res.scores, masks=res.masks)

    # ---- mmdet/rsprompter/models.py:661-715 RSMaskFormerFusionHead.predict (crop to the resized image + rescale)
    rf = load_defs("mmdet/rsprompter/models.py", ["predict"], cls="RSMaskFormerFusionHead")
    meta = dict(img_shape=(60, 96), ori_shape=(50, 80), scale_factor=(1.2, 1.2), batch_input_shape=(96, 96))
    self_ = SimpleNamespace(test_cfg=dict(panoptic_on=False, semantic_on=False, instance_on=True, max_per_image=7),
                            num_classes=ncls, num_things_classes=ncls)
    self_.instance_postprocess = lambda c, m: fh["instance_postprocess"](self_, c, m)
    mp_b = torch.randn(1, nq, 96, 96, generator=g) * 2
    out = rf["predict"](self_, mask_cls[None], mp_b, [SimpleNamespace(metainfo=meta)], rescale=True)[0]["ins_results"]
    fx["fusion_predict_rescale"] = dict(mask_cls=mask_cls, mask_pred=mp_b[0], meta=meta, num_classes=ncls, max_per_image=7,
                                        bboxes=out.bboxes, labels=out.labels, scores=out.scores, masks=out.masks)

    # ---- mmdet/rsprompter/models.py:1746-1784 RSPrompterAnchorMaskHead._predict_by_feat_single (rescale=True)
    mh = load_defs("mmdet/rsprompter/models.py", ["_predict_by_feat_single"], cls="RSPrompterAnchorMaskHead")
    meta2 = dict(ori_shape=(50, 80), scale_factor=(1.2, 1.2), batch_input_shape=(96, 96))
    logits = torch.randn(5, 1, 24, 24, generator=g) * 3
    boxes = torch.rand(5, 4, generator=g) * 90
    b_in = boxes.clone()
    im = mh["_predict_by_feat_single"](SimpleNamespace(), logits, b_in, torch.zeros(5, dtype=torch.long), meta2,
                                       SimpleNamespace(mask_thr_binary=0.5), rescale=True)
    fx["anchor_mask_rescale"] = dict(logits=logits, boxes=boxes, meta=meta2, masks=im, boxes_out=b_in)

    # ---- mmpretrain checkpoint resize at load time (VS:611-662, mmpretrain/models/utils/embed.py:16-59)
    import types
    fake_log = types.ModuleType("mmengine.logging")
    fake_log.MMLogger = SimpleNamespace(get_current_instance=lambda: SimpleNamespace(info=lambda *a, **k: None))
    sys.modules.setdefault("mmengine", types.ModuleType("mmengine"))
    sys.modules["mmengine.logging"] = fake_log
    em = load_defs("mmpretrain/models/utils/embed.py", ["resize_pos_embed"])
    vsl = load_defs("mmpretrain/models/backbones/vit_sam.py", ["_prepare_pos_embed", "_prepare_relative_position"],
                    cls="ViTSAM")
    vsl["_prepare_pos_embed"].__globals__.update(resize_pos_embed=em["resize_pos_embed"])
    ck_pos = torch.randn(1, 8, 8, 12, generator=g)
    ck_rel = torch.randn(15, 6, generator=g)
    ck_win = torch.randn(5, 6, generator=g)
    own = {"layers.0.attn.rel_pos_h": torch.zeros(7, 6), "layers.1.attn.rel_pos_w": torch.zeros(5, 6)}
    self_ = SimpleNamespace(pos_embed=torch.zeros(1, 4, 4, 12), patch_embed=SimpleNamespace(init_out_size=(4, 4)),
                            interpolate_mode="bicubic", embed_dims=12, state_dict=lambda: own)
    sd_ck = {"pos_embed": ck_pos.clone(), "layers.0.attn.rel_pos_h": ck_rel.clone(), "layers.1.attn.rel_pos_w": ck_win.clone()}
    vsl["_prepare_pos_embed"](self_, sd_ck, "")
    vsl["_prepare_relative_position"](self_, sd_ck, "")
    fx["ckpt_resize"] = dict(pos_embed=ck_pos, rel_pos=ck_rel, rel_win=ck_win, pos_embed_out=sd_ck["pos_embed"],
                             rel_pos_out=sd_ck["layers.0.attn.rel_pos_h"], rel_win_out=sd_ck["layers.1.attn.rel_pos_w"])

    torch.save(fx, OUT)
    print("wrote", OUT, {k: list(v.keys()) for k, v in fx.items()})


def main_maskrcnn():
    """tests/golden/reference_maskrcnn.pt: FCNMaskHead's paste path (SAMSegMaskRCNN), a file of its own so that the
    first fixture file stays byte-identical."""
    g = torch.Generator().manual_seed(4321)
    fx = {}
    rel = "mmdet/models/roi_heads/mask_heads/fcn_mask_head.py"
    pm = load_defs(rel, ["_do_paste_mask"])
    n, hm, H, W = 9, 28, 72, 100
    probs = torch.rand(n, 1, hm, hm, generator=g)
    boxes = torch.rand(n, 4, generator=g) * torch.tensor([60., 40., 60., 40.])
    boxes[:, 2:] = boxes[:, :2] + torch.rand(n, 2, generator=g) * 50 + 0.5
    boxes[0] = torch.tensor([-8.0, -3.5, 30.0, 90.0])          # sticks out of the canvas
    boxes[1] = torch.tensor([40.0, 10.0, 40.0, 30.0])          # zero width: inf -> 0 rule
    boxes[2] = torch.tensor([0.0, 0.0, float(W), float(H)])    # the whole canvas
    out, _ = pm["_do_paste_mask"](probs, boxes, H, W, skip_empty=False)
    fx["do_paste_mask"] = dict(probs=probs, boxes=boxes, hw=(H, W), out=out)

    mh = load_defs(rel, ["_predict_by_feat_single"], cls="FCNMaskHead")
    mh["_predict_by_feat_single"].__globals__.update(_do_paste_mask=pm["_do_paste_mask"], BYTES_PER_FLOAT=4,
                                                     GPU_MEM_LIMIT=1024 ** 3)
    C = 3
    logits = torch.randn(n, C, hm, hm, generator=g) * 3
    labels = torch.randint(0, C, (n,), generator=g)
    self_ = SimpleNamespace(class_agnostic=False)
    for name, rescale in (("fcn_predict_rescale", True), ("fcn_predict_norescale", False)):
        meta = dict(ori_shape=(60, 84), scale_factor=(1.25, 1.2))
        b_in = boxes.clone()
        im = mh["_predict_by_feat_single"](self_, logits.clone(), b_in, labels, meta,
                                           SimpleNamespace(mask_thr_binary=0.5), rescale=rescale)
        fx[name] = dict(logits=logits, boxes=boxes, labels=labels, meta=meta, rescale=rescale, masks=im, boxes_out=b_in)
    out_path = os.path.join(os.path.dirname(OUT), "reference_maskrcnn.pt")
    torch.save(fx, out_path)
    print("wrote", out_path, {k: list(v.keys()) for k, v in fx.items()})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("/root/reference not mounted")
    if "maskrcnn" in sys.argv[1:]:
        main_maskrcnn()
    else:
        main()
