"""A few whole-pipeline steps of one detector variant, for ncu launch lists / traffic captures and quick timings:
    python profiles/run_step.py --variant anchor|query|maskrcnn|mask2former [--arch base] [--batch 8] [--size 1024] [--steps 2]
Prints the CUDA-event time per step (meaningless under ncu) and the number of library launches."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="anchor", choices=["anchor", "query", "maskrcnn", "mask2former"])
    ap.add_argument("--arch", default="base")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--classes", type=int, default=10)
    a = ap.parse_args()
    from rsprompter_b200 import _lib, model_configs, synthetic
    from rsprompter_b200.model_configs import SELECT_LAYERS
    from rsprompter_b200.registry import MODELS
    from rsprompter_b200.sam_config import VISION_ARCHS
    arch = VISION_ARCHS[a.arch]
    nsel = len(SELECT_LAYERS[a.arch])
    if a.variant == "anchor":
        model = MODELS.build(model_configs.anchor_model_cfg(a.arch, a.classes))
        model.load_state_dict(synthetic.anchor_detector_state_dict(arch, a.classes, nsel, seed=0))
    elif a.variant == "maskrcnn":
        model = MODELS.build(model_configs.maskrcnn_model_cfg(a.arch, a.classes))
        model.load_state_dict(synthetic.maskrcnn_detector_state_dict(arch, a.classes, nsel, seed=0))
    elif a.variant == "mask2former":
        model = MODELS.build(model_configs.mask2former_model_cfg(a.arch, a.classes))
        model.load_state_dict(synthetic.mask2former_detector_state_dict(arch, a.classes, nsel, seed=0))
    else:
        model = MODELS.build(model_configs.query_model_cfg(a.arch, a.classes))
        model.load_state_dict(synthetic.query_detector_state_dict(arch, a.classes, nsel, seed=0))
    model = model.cuda()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(a.batch, 3, a.size, a.size, generator=g).cuda()
    for _ in range(a.warmup):
        model.predict(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count
    e0.record()
    for _ in range(a.steps):
        model.predict(x)
    e1.record()
    torch.cuda.synchronize()
    launches = (_lib.launch_count - l0) // a.steps
    # host-side issue time of one step (no synchronisation inside): tells whether the step is launch-bound
    import time
    cpu_ms = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.predict_raw(x)
        cpu_ms.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
    print(json.dumps(dict(host_issue_ms_per_step=cpu_ms)))
    print(json.dumps(dict(variant=a.variant, arch=a.arch, batch=a.batch, size=a.size, steps=a.steps,
                          ms_per_step=e0.elapsed_time(e1) / a.steps, launches_per_step=launches,
                          images_per_s=a.batch * a.steps / (e0.elapsed_time(e1) * 1e-3))))


if __name__ == "__main__":
    main()
