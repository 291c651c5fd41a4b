#!/usr/bin/env python
"""bench.py -- RSPrompter inference hot path on B200 (contract: see the task brief / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the CPU oracle (reference arm)

Workload at N GPUs: BASELINE.json configs[1] -- RSPrompter-anchor ViT-B, bf16, batch 8 per GPU,
1024x1024 synthetic images, random-init weights of the exact architecture (seeded).  A step is one
full predict() of one batch: SAM ViT-B encoder -> RSFPN -> RPN -> RoI head -> mask head -> SAM mask
decoder -> sigmoid/bilinear/threshold masks.  One process per GPU, batch-sharded (weak scaling), one
all-gather of the per-image result records per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

ARCH, NUM_CLASSES, BATCH, SIZE = "base", 10, 8, 1024
N_INPUT_SETS = 3   # distinct input batches rotated through the timed loop (3 x 100 MB > 126 MB L2)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))),
                    hbm=float(p.get("hbm_gbs", 6650.0)), source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md: 1.4 PF sustained, 6.65 TB/s)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(float(parts[0]))
                    self.max_mhz = float(parts[1])
                    for nm, v in zip(names, parts[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        return dict(sm_mhz=statistics.median(self.samples) if self.samples else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(self.samples))


def _workload_config(n_gpus: int) -> dict:
    cfg = dict(workload=f"RSPrompter-anchor ViT-B bf16, bs={BATCH}/GPU, {SIZE}x{SIZE} synthetic (BASELINE.json configs[1])",
                num_classes=NUM_CLASSES, global_batch=BATCH * n_gpus, parallelism=f"dp{n_gpus} (batch-sharded)",
                l2_policy=f"{N_INPUT_SETS} distinct input batches rotated (> L2); activations per step >> L2",
                weights="seeded random init of the exact architecture")
    return cfg


def _flops_model():
    """Algorithmic FLOP of one GEMM launch = 2*M*N*K; summed over the tcgen05 GEMM launches of a step."""
    return None


# ================================================================================================
# reference arm / CPU baseline: the oracle port on the host cores
# ================================================================================================
def _oracle_setup():
    from oracle import restate_anchor as ra
    from rsprompter_b200 import synthetic
    from rsprompter_b200.model_configs import SELECT_LAYERS
    from rsprompter_b200.sam_config import VISION_ARCHS, SamDecoderArch
    arch = VISION_ARCHS[ARCH]
    sd = synthetic.anchor_detector_state_dict(arch, NUM_CLASSES, len(SELECT_LAYERS[ARCH]), seed=0)

    def run(x):
        with torch.no_grad():
            return ra.anchor_predict(sd, arch, SamDecoderArch(), x, NUM_CLASSES, SELECT_LAYERS[ARCH])
    return run


def cpu_baseline(max_seconds: float = 45.0) -> dict:
    """Oracle ('port') on the host cores, bounded sample: whole pipeline on 1 image (bs=1) of the workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run = _oracle_setup()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, SIZE, SIZE, generator=g)
    t0 = time.perf_counter()
    run(x)
    dt = time.perf_counter() - t0
    n = 1
    if dt < max_seconds / 3:   # one more for a steadier number if cheap
        t0 = time.perf_counter()
        run(x)
        dt = time.perf_counter() - t0
    return dict(value=n / dt, unit="images/s", cores=cores, kind="port",
                sample=f"full anchor pipeline (oracle/restate*.py, fp32, torch {torch.__version__}) on 1 image "
                       f"{SIZE}x{SIZE}, bs=1, {dt:.1f} s")


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run = _oracle_setup()
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(1, 3, SIZE, SIZE, generator=g) for _ in range(2)]
    budget = 240.0
    t_start = time.perf_counter()
    w_done = 0
    for i in range(args.warmup):
        run(xs[i % 2])
        w_done += 1
        if time.perf_counter() - t_start > budget * 0.3:
            break
    times = []
    for i in range(args.steps):
        t0 = time.perf_counter()
        run(xs[i % 2])
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget:
            break
    total = sum(times)
    val = len(times) / total
    line = dict(metric="images/sec", value=val, unit="images/s", n_gpus=args.gpus, steps=len(times),
                warmup=w_done, ms_per_step=1e3 * total / len(times), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=_workload_config(args.gpus),
                cpu_baseline=dict(value=val, unit="images/s", cores=cores, kind="port",
                                  sample=f"each step = full anchor pipeline on 1 image (bs=1) of the workload; "
                                         f"{len(times)} of {args.steps} requested steps fit the {budget:.0f} s budget"),
                e2e=dict(value=val, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ================================================================================================
# this repo's arm
# ================================================================================================
def run_ours(args) -> None:
    import torch.distributed as dist
    from rsprompter_b200 import _lib, model_configs, synthetic
    from rsprompter_b200.model_configs import SELECT_LAYERS
    from rsprompter_b200.registry import MODELS, make_data_samples
    from rsprompter_b200.results import gather_mask_logits, gather_records, pack_records
    from rsprompter_b200.sam_config import VISION_ARCHS

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (the CUDA path has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    arch = VISION_ARCHS[ARCH]
    model = MODELS.build(model_configs.anchor_model_cfg(ARCH, NUM_CLASSES))
    model.load_state_dict(synthetic.anchor_detector_state_dict(arch, NUM_CLASSES, len(SELECT_LAYERS[ARCH]), seed=0))
    model = model.to(dev)
    g = torch.Generator().manual_seed(1000 + rank)
    host = [torch.randn(BATCH, 3, SIZE, SIZE, generator=g).pin_memory() for _ in range(N_INPUT_SETS)]
    resident = [h.to(dev) for h in host]
    samples = make_data_samples(BATCH, SIZE)
    thr = 0.5
    M = 100

    gather_masks = bool(os.environ.get("RSP_BENCH_GATHER_MASKS"))
    use_graph = not os.environ.get("RSP_BENCH_NO_GRAPH")
    if use_graph:
        model.enable_cuda_graphs()      # the device-resident forward is captured once per input shape and replayed
    forward = model._raw if use_graph else model.predict_raw

    def step_resident(i):
        r = forward(resident[i % N_INPUT_SETS])
        masks = _lib.mask_paste(r["mask_logits"][:, 0].contiguous(), (SIZE, SIZE), thr, 0)
        rec = pack_records(r["bboxes"], r["scores"], r["labels"])      # [B, M, 6]
        rec, cnt = gather_records(rec, r["counts"])                    # the one collective of the path
        if gather_masks:                                               # opt-in: fp16 256^2 logits ride along
            gather_mask_logits(r["mask_logits"][:, 0])
        return masks, rec, cnt

    def step_e2e(i):
        x = host[i % N_INPUT_SETS] if use_graph else host[i % N_INPUT_SETS].to(dev, non_blocking=True)
        r = forward(x)                  # graph path: the pinned batch is copied straight into the static input
        masks = _lib.mask_paste(r["mask_logits"][:, 0].contiguous(), (SIZE, SIZE), thr, 0)
        rec = pack_records(r["bboxes"], r["scores"], r["labels"])
        gather_records(rec, r["counts"])
        rec_h = rec.cpu()
        cnt_h = r["counts"].cpu()
        return masks, rec_h, cnt_h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    for i in range(2):
        step_e2e(i)
    barrier()

    with ClockSampler(local) as clk:
        l0 = _lib.launch_count
        ms = timed(step_resident, args.steps)
        launches = _lib.launch_count - l0
    ms_e2e = timed(step_e2e, args.steps)

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): an instrumented pass of the same steps,
    # CUDA events around every GEMM launch on the launching stream
    records = []
    orig_gemm = _lib.gemm

    def gemm_timed(a, w, *pa, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_gemm(a, w, *pa, **kw)
        e1.record()
        if not kw.get("simt"):
            records.append((e0, e1, 2.0 * a.shape[0] * w.shape[0] * a.shape[1]))
        return out

    peaks = _peaks()
    roof = None
    if rank == 0:
        _lib.gemm = gemm_timed
        try:
            inst_steps = min(args.steps, 3)
            torch.cuda.synchronize()
            for i in range(inst_steps):
                # park the GPU (~50 ms) so the host enqueues the whole instrumented step ahead of it: each event
                # pair then brackets exactly one GEMM on a never-starved stream (without this the eager, event-laden
                # pass is host-bound and the pairs would include launch gaps)
                park = getattr(torch.cuda, "_sleep", None)
                if park is not None:
                    park(100_000_000)
                model.predict_raw(resident[i % N_INPUT_SETS])
                torch.cuda.synchronize()
        finally:
            _lib.gemm = orig_gemm
        t_ms = sum(a.elapsed_time(b) for a, b, _ in records)
        fl = sum(f for _, _, f in records)
        ach = fl / (t_ms * 1e-3) / 1e12 if t_ms > 0 else 0.0
        traffic, traffic_src = None, None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_traffic_anchor_vitb.json")
        if os.path.exists(tpath):      # ncu dram__bytes_read.sum + dram__bytes_write.sum of the same step, per launch
            with open(tpath) as f:
                tj = json.load(f)
            g = [k for k in tj["kernels"] if "gemm_bf16_tcgen05" in k["kernel"]]
            if g:
                traffic = sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in g) / sum(k["launches"] for k in g)
                traffic_src = "profiles/r01_traffic_anchor_vitb.json (ncu capture of one step, mean bytes per GEMM launch)"
        roof = dict(bound="tensor", kernel="gemm_bf16_tcgen05_kernel (all tile shapes)", achieved=ach,
                    peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"], traffic=traffic,
                    traffic_source=traffic_src,
                    peak_source=peaks["source"], launches_per_step=len(records) // max(inst_steps, 1),
                    gemm_ms_per_step=t_ms / max(inst_steps, 1),
                    algorithmic_tflop_per_step=fl / max(inst_steps, 1) / 1e12,
                    note="achieved = sum(2*M*N*K) / sum(CUDA-event time) over every tcgen05 GEMM launch of the step")
    if world > 1:
        dist.barrier()

    if rank == 0:
        imgs = BATCH * world * args.steps
        value = imgs / (ms * 1e-3)
        e2e_val = imgs / (ms_e2e * 1e-3)
        h2d = BATCH * 3 * SIZE * SIZE * 4
        d2h = BATCH * M * 6 * 4 + BATCH * 4
        cb = cpu_baseline() if (world == 1 and not os.environ.get("RSP_BENCH_SKIP_CPU")) else None
        line = dict(metric="images/sec", value=value, unit="images/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                    config=_workload_config(world), clocks=clk.summary(), cuda_graph=use_graph,
                    e2e=dict(value=e2e_val, unit="images/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             ms_per_step=ms_e2e / args.steps,
                             note="pinned host batch -> predict() -> detection records + counts read back; "
                                  "boolean masks stay on the device as in the reference's predict()"),
                    gpu_launches=launches, roofline=roof)
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
