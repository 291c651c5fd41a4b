"""Degenerate results through the fixed-size padded candidate lists: images with no detection at all (every score under
score_thr; StandardRoIHead.predict_bbox's empty branch, standard_roi_head.py:306-317) and images without a single RPN
proposal (every decoded box under min_bbox_size, rpn_head.py:267-271) must come out as empty instance sets - through
predict(), the rescale path and the result record - next to a normal image in the same batch state."""
import pytest
import torch

pytestmark = pytest.mark.gpu

NUM_CLASSES = 10


def _build(kind):
    from rsprompter_b200 import model_configs, synthetic
    from rsprompter_b200.registry import MODELS
    from rsprompter_b200.sam_config import VISION_ARCHS
    arch = VISION_ARCHS["base"]
    if kind == "anchor":
        m = MODELS.build(model_configs.anchor_model_cfg("base", NUM_CLASSES))
        m.load_state_dict(synthetic.anchor_detector_state_dict(arch, NUM_CLASSES, 6, seed=3), strict=True)
    else:
        m = MODELS.build(model_configs.maskrcnn_model_cfg("base", NUM_CLASSES))
        m.load_state_dict(synthetic.maskrcnn_detector_state_dict(arch, NUM_CLASSES, 6, seed=3), strict=True)
    return m.cuda()


def _assert_empty(p, hw):
    assert len(p) == 0 and tuple(p.bboxes.shape) == (0, 4) and tuple(p.scores.shape) == (0,)
    assert tuple(p.labels.shape) == (0,) and tuple(p.masks.shape) == (0, *hw) and p.masks.dtype == torch.bool


@pytest.mark.parametrize("kind", ["anchor", "maskrcnn"])
def test_no_detection_above_score_thr(kind):
    from rsprompter_b200.registry import make_data_samples
    m = _build(kind)
    torch.manual_seed(1)
    x = torch.randn(2, 3, 1024, 1024).cuda()
    assert len(m.predict(x)[0].pred_instances) > 0
    m.roi_head.test_cfg["score_thr"] = 1.5                     # softmax scores never exceed 1
    out = m.predict(x)
    for o in out:
        _assert_empty(o.pred_instances, (1024, 1024))
    ds = make_data_samples(2, 1024)
    ds[1].set_metainfo(dict(ori_shape=(600, 800), img_shape=(768, 1024), scale_factor=(1.28, 1.28)))
    out = m.predict(x, ds)
    _assert_empty(out[0].pred_instances, (1024, 1024))
    _assert_empty(out[1].pred_instances, (600, 800))
    rec = m.predict_records(x)
    assert rec.counts.tolist() == [0, 0]
    inst = rec.instances()
    assert inst[0]["masks"].shape[0] == 0 and inst[1]["bboxes"].shape == (0, 4)
    host = rec.to_host(non_blocking=False).instances()
    assert host[0]["masks"].shape[0] == 0


@pytest.mark.parametrize("kind", ["anchor", "maskrcnn"])
def test_no_rpn_proposal(kind):
    m = _build(kind)
    torch.manual_seed(2)
    x = torch.randn(1, 3, 1024, 1024).cuda()
    m.rpn_head.test_cfg["min_bbox_size"] = 5000                # no decoded box is that large after clipping to 1024
    raw = m.predict_raw(x)
    torch.cuda.synchronize()
    assert int(raw["counts"][0]) == 0
    assert torch.isfinite(raw["bboxes"]).all() and torch.isfinite(raw["scores"]).all()
    _assert_empty(m.predict(x)[0].pred_instances, (1024, 1024))


def test_query_detector_with_one_query_and_odd_batch():
    """RSPrompter-query with a single query (max_per_image = 1) on a batch of 3: the grouped mask-embed GEMM, the
    top-k and the record all run at their smallest sizes."""
    from rsprompter_b200 import model_configs, sam_config, synthetic
    from rsprompter_b200.registry import MODELS
    cfg = model_configs.query_model_cfg("base", NUM_CLASSES, prompt_shape=(1, 5))
    m = MODELS.build(cfg)
    m.load_state_dict(synthetic.query_detector_state_dict(sam_config.VISION_ARCHS["base"], NUM_CLASSES, 6, nq=1, seed=5))
    m = m.cuda()
    torch.manual_seed(3)
    x = torch.randn(3, 3, 1024, 1024).cuda()
    out = m.predict(x)
    rec = m.predict_records(x)
    torch.cuda.synchronize()
    assert rec.counts.tolist() == [1, 1, 1]
    for b in range(3):
        p = out[b].pred_instances
        assert len(p) == 1 and p.masks.shape == (1, 1024, 1024) and torch.isfinite(p.scores).all()
        i = rec.instances()[b]
        assert torch.equal(i["masks"], p.masks) and torch.equal(i["bboxes"], p.bboxes)
