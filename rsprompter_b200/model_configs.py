"""Programmatic model configs equivalent to the reference's ``configs/rsprompter`` files, for
benchmarks and tests on boxes where /root/reference is not mounted.  Field values follow
configs/rsprompter/_base_/rsprompter_anchor.py:57-200 and rsprompter_anchor-nwpu.py:19-65 /
rsprompter_anchor-nwpu-peft-512.py:59-101; checkpoints are omitted (random-init weights)."""
from __future__ import annotations

SELECT_LAYERS = {"base": range(1, 13, 2), "large": range(1, 25, 2), "huge": range(1, 33, 2)}


def anchor_model_cfg(arch: str = "base", num_classes: int = 10, points: int = 5, mmpretrain_img_size: int | None = None) -> dict:
    name = f"facebook/sam-vit-{arch}"
    if mmpretrain_img_size is None:
        backbone = dict(type="RSSamVisionEncoder", hf_pretrain_name=name,
                        extra_config=dict(output_hidden_states=True))
        aggregator = dict(type="RSFeatureAggregator", in_channels=name, out_channels=256, hidden_channels=32,
                          select_layers=SELECT_LAYERS[arch])
    else:
        backbone = dict(type="MMPretrainSamVisionEncoder", hf_pretrain_name=name, img_size=mmpretrain_img_size)
        aggregator = dict(type="PseudoFeatureAggregator", in_channels=256, hidden_channels=512, out_channels=256)
    return dict(
        type="RSPrompterAnchor",
        decoder_freeze=False,
        shared_image_embedding=dict(type="RSSamPositionalEmbedding", hf_pretrain_name=name),
        backbone=backbone,
        neck=dict(type="RSFPN", feature_aggregator=aggregator,
                  feature_spliter=dict(type="RSSimpleFPN", backbone_channel=256, in_channels=[64, 128, 256, 256],
                                       out_channels=256, num_outs=5, norm_cfg=dict(type="LN2d", requires_grad=True))),
        rpn_head=dict(type="RPNHead", in_channels=256, feat_channels=256,
                      anchor_generator=dict(type="AnchorGenerator", scales=[4, 8], ratios=[0.5, 1.0, 2.0],
                                            strides=[4, 8, 16, 32, 64]),
                      bbox_coder=dict(type="DeltaXYWHBBoxCoder", target_means=[0., 0., 0., 0.],
                                      target_stds=[1.0, 1.0, 1.0, 1.0])),
        roi_head=dict(
            type="RSPrompterAnchorRoIPromptHead", with_extra_pe=True,
            bbox_roi_extractor=dict(type="SingleRoIExtractor",
                                    roi_layer=dict(type="RoIAlign", output_size=7, sampling_ratio=0),
                                    out_channels=256, featmap_strides=[4, 8, 16, 32]),
            bbox_head=dict(type="Shared2FCBBoxHead", in_channels=256, fc_out_channels=1024, roi_feat_size=7,
                           num_classes=num_classes,
                           bbox_coder=dict(type="DeltaXYWHBBoxCoder", target_means=[0., 0., 0., 0.],
                                           target_stds=[0.1, 0.1, 0.2, 0.2]),
                           reg_class_agnostic=False),
            mask_roi_extractor=dict(type="SingleRoIExtractor",
                                    roi_layer=dict(type="RoIAlign", output_size=14, sampling_ratio=0),
                                    out_channels=256, featmap_strides=[4, 8, 16, 32]),
            mask_head=dict(type="RSPrompterAnchorMaskHead",
                           mask_decoder=dict(type="RSSamMaskDecoder", hf_pretrain_name=name),
                           in_channels=256, roi_feat_size=14, per_pointset_point=points, with_sincos=True,
                           multimask_output=False, class_agnostic=True)),
        test_cfg=dict(rpn=dict(nms_pre=1000, max_per_img=1000, nms=dict(type="nms", iou_threshold=0.7),
                               min_bbox_size=0),
                      rcnn=dict(score_thr=0.05, nms=dict(type="nms", iou_threshold=0.5), max_per_img=100,
                                mask_thr_binary=0.5)))


def _backbone_neck(arch: str, mmpretrain_img_size: int | None):
    name = f"facebook/sam-vit-{arch}"
    if mmpretrain_img_size is None:
        backbone = dict(type="RSSamVisionEncoder", hf_pretrain_name=name,
                        extra_config=dict(output_hidden_states=True))
        aggregator = dict(type="RSFeatureAggregator", in_channels=name, out_channels=256, hidden_channels=32,
                          select_layers=SELECT_LAYERS[arch])
    else:
        backbone = dict(type="MMPretrainSamVisionEncoder", hf_pretrain_name=name, img_size=mmpretrain_img_size)
        aggregator = dict(type="PseudoFeatureAggregator", in_channels=256, hidden_channels=512, out_channels=256)
    neck = dict(type="RSFPN", feature_aggregator=aggregator,
                feature_spliter=dict(type="RSSimpleFPN", backbone_channel=256, in_channels=[64, 128, 256, 256],
                                     out_channels=256, num_outs=5, norm_cfg=dict(type="LN2d", requires_grad=True)))
    return name, backbone, neck


def query_model_cfg(arch: str = "base", num_classes: int = 10, prompt_shape: tuple = (100, 5),
                    mmpretrain_img_size: int | None = None) -> dict:
    """configs/rsprompter/_base_/rsprompter_query.py:57-190 + rsprompter_query-nwpu.py overrides."""
    name, backbone, neck = _backbone_neck(arch, mmpretrain_img_size)
    attn = dict(embed_dims=128, num_heads=8, dropout=0.0, batch_first=True)
    ffn = dict(embed_dims=128, feedforward_channels=512, num_fcs=2, ffn_drop=0.0, act_cfg=dict(type="ReLU", inplace=True))
    return dict(
        type="RSPrompterQuery",
        decoder_freeze=False,
        shared_image_embedding=dict(type="RSSamPositionalEmbedding", hf_pretrain_name=name),
        backbone=backbone,
        neck=neck,
        panoptic_head=dict(
            type="RSMask2FormerHead", decoder_plus=True,
            mask_decoder=dict(type="RSSamMaskDecoder", hf_pretrain_name=name),
            per_pointset_point=prompt_shape[1], with_sincos=True, multimask_output=False,
            in_channels=[256, 256, 256, 256, 256], feat_channels=128, out_channels=256,
            num_things_classes=num_classes, num_stuff_classes=0, num_queries=prompt_shape[0],
            num_transformer_feat_level=3,
            pixel_decoder=dict(
                type="MSDeformAttnPixelDecoder", strides=[4, 8, 16, 32, 64], num_outs=3,
                norm_cfg=dict(type="GN", num_groups=32), act_cfg=dict(type="ReLU"),
                encoder=dict(num_layers=3, layer_cfg=dict(
                    self_attn_cfg=dict(embed_dims=128, num_heads=8, num_levels=3, num_points=4, dropout=0.0,
                                       batch_first=True), ffn_cfg=ffn)),
                positional_encoding=dict(num_feats=64, normalize=True)),
            enforce_decoder_input_project=False,
            positional_encoding=dict(num_feats=64, normalize=True),
            transformer_decoder=dict(return_intermediate=True, num_layers=6,
                                     layer_cfg=dict(self_attn_cfg=attn, cross_attn_cfg=attn, ffn_cfg=ffn),
                                     init_cfg=None)),
        panoptic_fusion_head=dict(type="RSMaskFormerFusionHead", num_things_classes=num_classes,
                                  num_stuff_classes=0, loss_panoptic=None, init_cfg=None),
        test_cfg=dict(panoptic_on=False, semantic_on=False, instance_on=True, max_per_image=prompt_shape[0],
                      iou_thr=0.8, filter_low_score=True))


def maskrcnn_model_cfg(arch: str = "base", num_classes: int = 10) -> dict:
    """configs/rsprompter/_base_/samseg-maskrcnn.py:57-184 + samseg-maskrcnn-nwpu.py overrides."""
    name, backbone, neck = _backbone_neck(arch, None)
    return dict(
        type="SAMSegMaskRCNN",
        backbone=backbone,
        neck=neck,
        rpn_head=dict(type="RPNHead", in_channels=256, feat_channels=256,
                      anchor_generator=dict(type="AnchorGenerator", scales=[8], ratios=[0.5, 1.0, 2.0],
                                            strides=[4, 8, 16, 32, 64]),
                      bbox_coder=dict(type="DeltaXYWHBBoxCoder", target_means=[0., 0., 0., 0.],
                                      target_stds=[1.0, 1.0, 1.0, 1.0])),
        roi_head=dict(
            type="StandardRoIHead",
            bbox_roi_extractor=dict(type="SingleRoIExtractor",
                                    roi_layer=dict(type="RoIAlign", output_size=7, sampling_ratio=0),
                                    out_channels=256, featmap_strides=[4, 8, 16, 32]),
            bbox_head=dict(type="Shared2FCBBoxHead", in_channels=256, fc_out_channels=1024, roi_feat_size=7,
                           num_classes=num_classes,
                           bbox_coder=dict(type="DeltaXYWHBBoxCoder", target_means=[0., 0., 0., 0.],
                                           target_stds=[0.1, 0.1, 0.2, 0.2]),
                           reg_class_agnostic=False),
            mask_roi_extractor=dict(type="SingleRoIExtractor",
                                    roi_layer=dict(type="RoIAlign", output_size=14, sampling_ratio=0),
                                    out_channels=256, featmap_strides=[4, 8, 16, 32]),
            mask_head=dict(type="FCNMaskHead", num_convs=4, in_channels=256, conv_out_channels=256,
                           num_classes=num_classes)),
        test_cfg=dict(rpn=dict(nms_pre=1000, max_per_img=1000, nms=dict(type="nms", iou_threshold=0.7),
                               min_bbox_size=0),
                      rcnn=dict(score_thr=0.05, nms=dict(type="nms", iou_threshold=0.5), max_per_img=100,
                                mask_thr_binary=0.5)))


def mask2former_model_cfg(arch: str = "base", num_classes: int = 10, num_queries: int = 100) -> dict:
    """configs/rsprompter/_base_/samseg-mask2former.py:60-191 + samseg-mask2former-nwpu.py overrides."""
    name, backbone, neck = _backbone_neck(arch, None)
    attn = dict(embed_dims=256, num_heads=8, dropout=0.0, batch_first=True)
    return dict(
        type="SAMSegMask2Former",
        backbone=backbone,
        neck=neck,
        panoptic_head=dict(
            type="Mask2FormerHead", in_channels=[256, 256, 256, 256, 256], feat_channels=256, out_channels=256,
            num_things_classes=num_classes, num_stuff_classes=0, num_queries=num_queries, num_transformer_feat_level=3,
            pixel_decoder=dict(
                type="MSDeformAttnPixelDecoder", strides=[4, 8, 16, 32, 64], num_outs=3,
                norm_cfg=dict(type="GN", num_groups=32), act_cfg=dict(type="ReLU"),
                encoder=dict(num_layers=3, layer_cfg=dict(
                    self_attn_cfg=dict(embed_dims=256, num_heads=8, num_levels=3, num_points=4, dropout=0.0,
                                       batch_first=True),
                    ffn_cfg=dict(embed_dims=256, feedforward_channels=1024, num_fcs=2, ffn_drop=0.0,
                                 act_cfg=dict(type="ReLU", inplace=True)))),
                positional_encoding=dict(num_feats=128, normalize=True)),
            enforce_decoder_input_project=False,
            positional_encoding=dict(num_feats=128, normalize=True),
            transformer_decoder=dict(return_intermediate=True, num_layers=9,
                                     layer_cfg=dict(self_attn_cfg=attn, cross_attn_cfg=attn,
                                                    ffn_cfg=dict(embed_dims=256, feedforward_channels=2048, num_fcs=2,
                                                                 ffn_drop=0.0, act_cfg=dict(type="ReLU", inplace=True))),
                                     init_cfg=None)),
        panoptic_fusion_head=dict(type="MaskFormerFusionHead", num_things_classes=num_classes, num_stuff_classes=0,
                                  loss_panoptic=None, init_cfg=None),
        test_cfg=dict(panoptic_on=False, semantic_on=False, instance_on=True, max_per_image=100, iou_thr=0.8,
                      filter_low_score=True))
