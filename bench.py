#!/usr/bin/env python
"""Benchmark of the detection hot path on B200 (contract: see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            # this framework, N ranks (torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A step = one pass of the hot path over one batch of 32 synthetic 416x416 frames:
Darknet-53 (yolov3.cfg) forward -> YOLO decode -> confidence filter + NMS, fp16 compute.
`value`  : frames/s with the input batch already resident in HBM.
`e2e`    : frames/s through the public API (millieye_b200.models.DetectPipeline: Darknet forward + NMS) with
           the batch in pinned HOST memory - H2D copy of the images and D2H read of the detections
           inside the timed region.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 32
SIZE = 416
CFG = "yolov3"
CONF_THRESH = 0.2     # test_fusion.py:143
CPU_BATCH = 4         # bounded CPU sample (Darknet-53 on CPU is run at N=4)
METRIC = "frames/sec at 416x416 batch32"
# Synthetic head statistics (random-init heads give conf ~ 0.5 everywhere): head logits of moderate width
# (std ~ 0.3: boxes near their anchor sizes, like a trained detector's) and an objectness bias that lets ~75 of the 10 647
# boxes per frame pass the 0.2 confidence filter, of which the NMS keeps ~18.  Wide logits (the round-1 recipe, gain 3) only
# amplify the fp16 error through exp() in the decode; narrower ones (gain 0.5) pile hundreds of boxes within 1e-3 of one
# confidence value, so that any threshold lands on a cliff and the kept set depends on differences below the tolerance.
WEIGHTS = dict(obj_bias=-2.0, head_gain=1.0)


# tiny-12 heads of BASELINE configs 3 and 4: ~180 of 2535 boxes per frame pass the 0.2 filter, the NMS keeps ~80 of which
# ~23 are class 0 (the fusion model's proposals), median box width ~270 px.  (Round 2's first recipe, bias -3 / gain 1, gave
# the same proposal count but box widths of thousands of pixels - RoIAlign's adaptive sampling grid then walks hundreds
# of samples per bin, a cost no real detector output has.)
FUSION_WEIGHTS = dict(obj_bias=-2.0, head_gain=0.3)

_JSON_FD = None


def quiet_stdout():
    """Keeps fd 1 for the JSON line only: libraries that write to stdout on their own (NCCL prints its version
    banner there on the first communicator) are sent to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def peaks():
    """Roofline denominators: the driver-measured MEASURED_PEAKS.json (burst = cuBLAS bf16 timed alone, sustained = the
    same GEMM looped for seconds under the power cap), else the profiling recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    src="measured")
    return dict(hbm_gbs=6650.0, burst=1590.0, sustained=1400.0, src="fallback")


def measured_traffic():
    """DRAM bytes (read + write) of one step's conv launches, from the committed ncu capture
    (profiles/round2/traffic.json, else round 1's; `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over this
    script).  The capture is per step, like `achieved`; None if the profile is not in the tree."""
    path = os.path.join(ROOT, "profiles", "round2", "traffic.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "round1", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        t = json.load(fh)
    return dict(dram_bytes_per_step=t["dram_read_bytes"] + t["dram_write_bytes"], read=t["dram_read_bytes"],
                write=t["dram_write_bytes"], launches=t["conv_launches"], source=t["source"])


def l2_feed():
    """What actually bounds the dominant kernel: L2 -> SM operand bytes per clock of the conv chain kernel against the
    L2 slices' throughput cap, from the committed ncu capture (profiles/round2/l2_feed.json).  None if not in the tree."""
    path = os.path.join(ROOT, "profiles", "round2", "l2_feed.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        d = json.load(fh)
    top = d["launches"][0]
    return dict(kernel=d["kernel"], bytes_per_clk=top["bytes_per_clk"], cap_bytes_per_clk=d["cap_bytes_per_clk"],
                frac=top["frac_of_lts_cap"], tensor_pipe_active_pct=top["tensor_pipe_active_pct"], source=d["source"],
                cap_source=d["cap_source"])


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as fh:
            for line in fh:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 6:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for name, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def build_models(device):
    from millieye_b200 import configs
    from millieye_b200.models import Darknet
    from oracle import synth  # weights recipe only (shared with the CPU baseline so both arms run the same net)
    net = Darknet(configs.cfg_path(CFG)).eval()
    sd = synth.fill_state_dict(net.state_dict(), seed=0, conv_gain=0.6, **WEIGHTS)
    net.load_state_dict(sd)
    net.to(device)
    return net, sd


def reference_step_fn(threads, batch):
    """The UNMODIFIED reference (baseline/_ref, staged by baseline/stage_reference.py) on the same step: its own
    Darknet.forward + non_max_suppression_cpp, same weights recipe and input distribution as the GPU arm."""
    from baseline import ref_runner
    from millieye_b200 import configs
    from millieye_b200.models import Darknet
    from oracle import synth
    sd = synth.fill_state_dict(Darknet(configs.cfg_path(CFG)).state_dict(), seed=0, conv_gain=0.6, **WEIGHTS)
    x = synth.synth_images(batch, SIZE, seed=0)
    return ref_runner.darknet53_step_fn(sd, x, CONF_THRESH, threads)


def reference_available():
    try:
        from baseline import ref_runner
        return ref_runner.available()
    except Exception:  # noqa: BLE001
        return False


def cpu_forward_fn(threads):
    """The reference's CPU path for the same step (oracle port: torch CPU fp32 ops + torchvision NMS); only used when
    the verbatim reference is not staged under baseline/_ref."""
    from millieye_b200 import configs
    from millieye_b200.models import Darknet
    from oracle import boxes as obox
    from oracle import darknet as odark
    from oracle import synth
    from oracle.parse_config import parse_model_config
    torch.set_num_threads(threads)
    md = parse_model_config(configs.cfg_path(CFG))
    sd = synth.fill_state_dict(Darknet(configs.cfg_path(CFG)).state_dict(), seed=0, conv_gain=0.6, **WEIGHTS)
    x = synth.synth_images(CPU_BATCH, SIZE, seed=0)

    def step():
        with torch.no_grad():
            _, y = odark.darknet_forward(md, sd, x)
            obox.non_max_suppression_cpp(y.clone().numpy(), CONF_THRESH, use_torchvision=True)
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    use_ref = reference_available() and not os.environ.get("ME_BENCH_PORT")
    if use_ref:
        # the reference itself at the full batch: one step = 32 frames, like the GPU arm
        batch, kind = BATCH, "reference"
        step = reference_step_fn(threads, batch)
        what = "the unmodified reference (baseline/_ref): Darknet.forward + non_max_suppression_cpp"
    else:
        batch, kind = CPU_BATCH, "port"
        step = cpu_forward_fn(threads)
        what = "oracle port (baseline/_ref not staged)"
    warm = min(args.warmup, 1 if use_ref else 2)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = batch * args.steps / dt
    sample = (f"{what}; Darknet-53 forward + decode + NMS, fp32, batch {batch} per step, {args.steps} steps, torch CPU ops "
              f"on {threads} threads")
    line = dict(impl="reference", metric=METRIC, value=fps, unit="frames/s", n_gpus=args.gpus, steps=args.steps,
                warmup=warm, ms_per_step=dt / args.steps * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"Darknet-53 YOLOv3 inference (forward + decode + conf filter + NMS), batch {BATCH} per "
                                     f"step, {SIZE}x{SIZE}", cpu_sample_batch=batch, same_config=batch == BATCH,
                            conf_thresh=CONF_THRESH),
                cpu_baseline=dict(value=fps, unit="frames/s", cores=threads, kind=kind, sample=sample),
                e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def match_detections(ref, det, cnt, frames, conf_thresh, tol=1e-3):
    """Row-by-row comparison of NMS survivors: ref[f] = oracle rows (k, 7+) or None, det (n, max_det, 7+) / cnt (n,) = the
    rows under test; columns x1, y1, x2, y2, conf, class_conf, class.  Rows are matched by class and IoU >= 0.9 (each
    oracle row once); a row that only one side reports must be explained by a score difference <= tol: its confidence
    is within tol of the threshold, it overlaps (IoU > 0.5 - tol) a same-class row of the other side whose score is
    within tol of its own (the NMS order of the two flipped), or it overlaps an unmatched row of the other side (the row
    that suppressed it there is itself such a flip)."""
    import numpy as np
    rows_ref = rows_gpu = matched = exact_order = 0
    near_threshold = near_tie = unexplained = 0
    box_err = score_err = 0.0

    def iou(a, b):
        ix = max(0.0, min(a[2], b[2]) - max(a[0], b[0]))
        iy = max(0.0, min(a[3], b[3]) - max(a[1], b[1]))
        inter = ix * iy
        u = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
        return inter / u if u > 0 else 0.0

    def explain(row, other, other_only):
        if abs(row[4] - conf_thresh) <= tol:
            return "threshold"
        for o in other:
            if o[6] == row[6] and abs(o[4] * o[5] - row[4] * row[5]) <= tol and iou(row, o) > 0.5 - tol:
                return "tie"
        for o in other_only:
            if o[6] == row[6] and iou(row, o) > 0.5 - tol:
                return "tie"
        return None

    for f in range(frames):
        r = ref[f] if ref[f] is not None else np.zeros((0, det.shape[2]), np.float32)
        r = np.asarray(r)
        g = det[f, :cnt[f]]
        rows_ref += len(r)
        rows_gpu += len(g)
        used = np.zeros(len(r), bool)
        got_used = np.zeros(len(g), bool)
        for i, row in enumerate(g):
            best, bj = 0.0, -1
            for j, rr in enumerate(r):
                if used[j] or rr[6] != row[6]:
                    continue
                v = iou(row, rr)
                if v > best:
                    best, bj = v, j
            if bj >= 0 and best >= 0.9:
                used[bj] = got_used[i] = True
                matched += 1
                exact_order += int(bj == i)
                rr = r[bj]
                scale = max(abs(rr[2] - rr[0]), abs(rr[3] - rr[1]), 1.0)
                box_err = max(box_err, float(np.abs(row[:4] - rr[:4]).max() / scale))
                score_err = max(score_err, float(np.abs(row[4:6] - rr[4:6]).max()))
        for rows, flags, other, oflags in ((g, got_used, r, used), (r, used, g, got_used)):
            for row in rows[~flags]:
                why = explain(row, other, other[~oflags])
                near_threshold += int(why == "threshold")
                near_tie += int(why == "tie")
                unexplained += int(why is None)
    return dict(checked=True, frames=frames, rows_oracle=rows_ref, rows_gpu=rows_gpu, rows_matched=matched,
                rows_same_rank=exact_order, rows_only_one_side=dict(conf_within_tol_of_threshold=near_threshold,
                                                                    nms_order_flip_within_tol=near_tie, unexplained=unexplained),
                max_box_err_rel=box_err, max_score_err_abs=score_err, tolerance=tol,
                within_tolerance=bool(unexplained == 0 and matched > 0 and box_err <= tol and score_err <= tol),
                how="oracle (CPU fp32 restatement of the reference) on the same images / weights; rows matched by class and "
                    "IoU >= 0.9; a row only one side reports must be explained by a score difference <= tolerance (at the "
                    "confidence threshold, or an NMS order flip between two overlapping rows)")


def check_parity(host_batch, rec, frames):
    """Detections of the first `frames` frames of a timed batch (as read back by the e2e arm) against the CPU oracle on
    the same images and weights (north_star: 1e-3 relative); see match_detections."""
    from millieye_b200 import configs
    from millieye_b200.models import Darknet
    from oracle import boxes as obox
    from oracle import darknet as odark
    from oracle import synth
    from oracle.parse_config import parse_model_config
    try:
        md = parse_model_config(configs.cfg_path(CFG))
        sd = synth.fill_state_dict(Darknet(configs.cfg_path(CFG)).state_dict(), seed=0, conv_gain=0.6, **WEIGHTS)
        x = host_batch[:frames].clone()
        if x.dtype == torch.uint8:
            x = x.float() / 255.0     # ToTensor, what the device applies to uploaded bytes
        with torch.no_grad():
            _, y = odark.darknet_forward(md, sd, x)
        ref, _ = obox.non_max_suppression_cpp(y.clone().numpy(), CONF_THRESH, use_torchvision=True)
        return match_detections(ref, rec.host_det.numpy(), rec.host_cnt.numpy(), frames, CONF_THRESH)
    except Exception as e:  # noqa: BLE001
        return dict(checked=False, error=str(e)[:200])


# ------------------------------------------------------------------------------------------------ secondary configs
def _dist_setup(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    return world, rank, local, device


def _timed_steps(fn, steps, warmup, world, device):
    """W warm-up steps, then K steps between CUDA events (barrier + synchronize on both sides), max over ranks (ms)."""
    import torch.distributed as dist
    for _ in range(max(warmup, 3)):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _hold_load(fn, ms_timed, steps, min_ms=700.0):
    """Keeps the same steps running (untimed) until about `min_ms` of load have passed, so that the 100 ms clock sampler
    sees the load of a timed region that only lasts tens of milliseconds.  The number of extra steps is derived from the
    max-over-ranks time, i.e. identical on every rank (the steps may contain collectives)."""
    per = max(ms_timed / max(steps, 1), 1e-3)
    for _ in range(max(0, int((min_ms - ms_timed) / per))):
        fn()
    torch.cuda.synchronize()


def _fusion_inputs(n, device, seed):
    """Synthetic frames + 64-point radar clouds: images U[0,1), points drawn from the fixture's
    empirical ranges, 0-3 radar boxes per frame."""
    import numpy as np
    from millieye_b200 import radar
    from oracle import synth
    gen = torch.Generator(device="cpu").manual_seed(seed)
    imgs_u8 = torch.randint(0, 256, (n, 3, SIZE, SIZE), generator=gen, dtype=torch.uint8)   # frames as bytes (cv2 / camera)
    imgs = imgs_u8.float() / 255.0
    rng = np.random.RandomState(seed)
    pts = np.stack([rng.uniform(-3, 3, (n, 64)), rng.uniform(1, 10, (n, 64)), rng.uniform(-1.5, 1.5, (n, 64)),
                    rng.uniform(-3, 3, (n, 64))], -1).astype(np.float32)
    return dict(imgs=imgs_u8.pin_memory(), imgs_dev=imgs.to(device), pts=torch.from_numpy(pts).to(device),
                cnt=torch.full((n,), 64, dtype=torch.int32, device=device), cfg=radar.make_cfg(out_size=SIZE // 16),
                boxes=synth.synth_radar_boxes(n, seed=seed).to(device))


def run_fusion(args):
    """BASELINE config 3: milliEye fusion (YOLOv3-tiny-12 + R-CNN refinement + radar MLP) inference, batch 32 per GPU,
    64 radar points per frame: radar heat-maps on the device, Network.forward, result rows on the host in the e2e arm."""
    import torch.distributed as dist
    from millieye_b200 import configs, radar
    from millieye_b200.engine import capture_graph
    from millieye_b200.my_models import FusionPipeline, Network, define_yolo
    from oracle import darknet as odark
    from oracle import synth
    from oracle.parse_config import parse_model_config
    world, rank, local, device = _dist_setup(args)
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=CONF_THRESH).eval()
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=0, **FUSION_WEIGHTS))
    model.to(device)
    inp = _fusion_inputs(BATCH, device, 100 + rank)
    last = {}

    pipe = FusionPipeline(model)

    def step_call():     # the reference scripts' own call pattern: one blocking forward per batch
        maps = radar.radar_maps(inp["pts"], inp["cnt"], inp["cfg"])
        last["out"] = model(inp["imgs_dev"], maps, inp["boxes"].clone(), 0)

    plan_b = model.base_detector.plan_for(BATCH, SIZE, device)
    for _ in range(2):   # frames resident in HBM before the timed region: both input slots of the detector's plan hold the
        plan_b.next_input().copy_(inp["imgs_dev"])   # batch (DarknetPlan.next_input(): the hand-over point for on-device
        model.base_detector.forward_device(plan_b.next_input())   # producers - submitting it copies nothing)
    torch.cuda.synchronize()

    def step_dev():      # stream of batches, frames resident in HBM, rows stay on the device
        maps = radar.radar_maps(inp["pts"], inp["cnt"], inp["cfg"])
        last["rec_dev"] = pipe.submit(plan_b.next_input(), maps, inp["boxes"].clone(), 0, readback=False)

    def step_e2e():      # stream of batches, uint8 frames from pinned host memory, every batch's rows read on the host
        maps = radar.radar_maps(inp["pts"], inp["cnt"], inp["cfg"])
        rec = pipe.submit(inp["imgs"], maps, inp["boxes"].clone(), 0, readback=True)
        prev = last.get("rec")
        if prev is not None:
            last["host"] = prev.wait()       # consume the previous batch while this one runs
        last["rec"] = rec

    for _ in range(3 * pipe.depth):      # every plan of the ring: first use, graph capture, first replay
        step_dev()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev = _timed_steps(step_dev, args.steps, args.warmup, world, device)
    _hold_load(step_dev, ms_dev, args.steps)
    clocks = sampler.stop() if sampler else None
    dev_rows = last["rec_dev"].wait().cpu()
    ms_e2e = _timed_steps(step_e2e, args.steps, args.warmup, world, device)
    last["host"] = last["rec"].wait().clone()
    d2h_bytes = int(last["rec"].host_flat.numel() * 4)
    ms_call = _timed_steps(step_call, args.steps, args.warmup, world, device)
    # the pipeline's rows are the blocking forward's rows (same inputs every step)
    p0 = [p for k, p in model._plans.items() if k[-1] == 0][0]       # the blocking forward's plan
    proposals = dict(image=int(p0.counts[0]), image_plus_radar=int(p0.counts[1]))
    pipeline_equals_forward = bool(torch.equal(dev_rows, last["out"].cpu()) and torch.equal(last["host"], last["out"].cpu()))
    # detector conv stack alone (HBM-bound on tiny-12): graph of the conv launches
    plan = model.base_detector.plan_for(BATCH, SIZE, device)
    conv_ops = {i for i, k in enumerate(plan.op_kinds) if k in ("conv", "maxpool", "upsample")}
    g = capture_graph(lambda: plan.enqueue_split(only=conv_ops))
    conv_ms = _timed_steps(g.replay, 20, 3, 1, device) / 20
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    md = parse_model_config(configs.cfg_path("yolov3-tiny-12"))
    act_bytes, w_bytes = odark.min_traffic_bytes(md, SIZE) if hasattr(odark, "min_traffic_bytes") else (31.4e6, 17.4e6)
    bytes_step = act_bytes * BATCH + w_bytes
    achieved = bytes_step / (conv_ms * 1e-3) / 1e9
    fps, fps_e2e = world * BATCH * args.steps / ms_dev * 1e3, world * BATCH * args.steps / ms_e2e * 1e3
    emit(dict(metric="frames/sec at 416x416 batch32, milliEye fusion (YOLO + R-CNN + radar MLP)", value=fps, unit="frames/s",
              n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_dev / args.steps, higher_is_better=True,
              scaling="weak", vs_baseline=None, dtype="f16", data="synthetic",
              config=dict(workload=f"milliEye fusion inference (BASELINE config 3): YOLOv3-tiny-12 detector + NMS + proposals + "
                                   f"score-map CNNs + PS-RoIAlign / RoIAlign + refinement / ensemble heads, batch {BATCH} per GPU, "
                                   f"{SIZE}x{SIZE}, 64 radar points per frame -> heat-maps on the device",
                          conf_thresh=CONF_THRESH, rows_last_step=int(last["out"].shape[0]),
                          proposals_last_step=proposals, synthetic_heads=FUSION_WEIGHTS,
                          api="FusionPipeline.submit per batch (proposal / head kernels of batch i on a second stream under the "
                              "backbone of batch i+1); blocking_forward_ms_per_step is one Network.forward call per batch",
                          blocking_forward_ms_per_step=ms_call / args.steps,
                          blocking_forward_fps=world * BATCH * args.steps / ms_call * 1e3,
                          pipeline_equals_forward=pipeline_equals_forward,
                          l2="~1 GB of activations per step exceed the 126 MB L2; no explicit flush"),
              clocks=clocks,
              e2e=dict(value=fps_e2e, unit="frames/s", h2d_bytes_per_step=BATCH * 3 * SIZE * SIZE,
                       d2h_bytes_per_step=d2h_bytes, ms_per_step=ms_e2e / args.steps,
                       output="rows + count of every batch in one copy to pinned host memory (the record's buffer)",
                       input="uint8 (N,3,S,S) frames in pinned host memory, ToTensor (x / 255) on the device"),
              gpu_launches=int((len(plan.ops) + len(plan.post_ops) + 14) * args.steps),
              roofline=dict(bound="hbm", achieved=achieved, peak=pk["hbm_gbs"], unit="GB/s", frac=achieved / pk["hbm_gbs"],
                            traffic=None, peak_source=pk["src"], kernel="tiny-12 conv stack (conv_gemm / conv_thin / conv_first_tc / "
                            "conv_chain / maxpool)", conv_ms_per_step=conv_ms,
                            note="achieved = minimum activation + weight bytes of the tiny-12 conv stack (31.4 MB / frame + 17.4 MB, "
                                 "SURVEY.md 8d) / device time of its launches (graph replay)"),
              cpu_baseline=None))
    if world > 1:
        dist.destroy_process_group()


def run_train3(args):
    """BASELINE config 4: stage-3 training step (fusion heads train, YOLO frozen), GLOBAL batch 64 split over the ranks,
    synchronised BatchNorm statistics + one gradient all-reduce over NCCL + the flat Adam kernel (train.py:169-191)."""
    import torch.distributed as dist
    from millieye_b200 import configs, radar
    from millieye_b200.my_models import Network, define_yolo
    from millieye_b200.stage3_train import Stage3Optimizer
    from oracle import synth
    world, rank, local, device = _dist_setup(args)
    GLOBAL = 64
    if GLOBAL % world:
        raise SystemExit("the global batch of 64 must divide over the ranks")
    per = GLOBAL // world
    model = Network(define_yolo(configs.cfg_path("yolov3-tiny-12")), conf_thresh=0.01)       # train.py:61
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), seed=0, **FUSION_WEIGHTS))
    model.to(device)
    inp = _fusion_inputs(per, device, 200 + rank)
    # ground truth near the detector's own proposals so that every label class occurs (positives, ignored, negatives)
    model.eval()
    maps0 = radar.radar_maps(inp["pts"], inp["cnt"], inp["cfg"])
    rows = model(inp["imgs_dev"], maps0, inp["boxes"].clone(), 1).cpu()
    tg = []
    for k, b in enumerate(rows[::9]):
        sh = (0.0, 0.03, 0.25)[k % 3]
        w, h = float(b[3] - b[1]), float(b[4] - b[2])
        x1, y1, x2, y2 = float(b[1]) + sh * w, float(b[2]) + sh * h, float(b[3]) + sh * w, float(b[4]) + sh * h
        tg.append([float(b[0]), 0.0, (x1 + x2) / 2 / SIZE, (y1 + y2) / 2 / SIZE, (x2 - x1) / SIZE, (y2 - y1) / SIZE])
    targets = torch.tensor(tg if tg else [[0, 0, 0.5, 0.5, 0.2, 0.2]], dtype=torch.float32)
    model.train()
    model.base_detector.eval()
    opt = Stage3Optimizer(model, lr=5e-4)
    last = {}

    def step(host_images):
        maps = radar.radar_maps(inp["pts"], inp["cnt"], inp["cfg"])
        loss, out, metric, att = model(inp["imgs"] if host_images else inp["imgs_dev"], maps, inp["boxes"].clone(), 0,
                                       targets.clone())
        model.backward_into(opt.grads)
        opt.all_reduce()
        opt.step()
        last["loss"] = float(loss) if host_images else loss
        last["rows"] = metric["total"]

    import random
    random.seed(1234 + rank)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev = _timed_steps(lambda: step(False), args.steps, args.warmup, world, device)
    _hold_load(lambda: step(False), ms_dev, args.steps)
    clocks = sampler.stop() if sampler else None
    ms_e2e = _timed_steps(lambda: step(True), args.steps, args.warmup, world, device)
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    fps, fps_e2e = GLOBAL * args.steps / ms_dev * 1e3, GLOBAL * args.steps / ms_e2e * 1e3
    bytes_step = 31.4e6 * per + 17.4e6
    emit(dict(metric="frames/sec, stage-3 training step, global batch 64", value=fps, unit="frames/s", n_gpus=world,
              steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_dev / args.steps, higher_is_better=True,
              scaling="strong", vs_baseline=None, dtype="f32 heads / f16 frozen detector", data="synthetic",
              config=dict(workload=f"stage-3 training step (BASELINE config 4): frozen YOLOv3-tiny-12 forward + NMS + proposals, "
                                   f"train-mode heads in fp32 (batch-statistics BatchNorm synchronised over ranks), labels + balanced "
                                   f"sample + focal / BCE loss, hand-derived backward of the {opt.numel} head parameters, one NCCL "
                                   f"all-reduce of the flat gradient buffer, Adam; global batch {GLOBAL} = {per} frames x {world} GPU(s), "
                                   f"{SIZE}x{SIZE}, 64 radar points per frame", proposals_per_rank_last_step=int(last["rows"]),
                          trainable_parameters=int(opt.numel), sync_bn=world > 1),
              clocks=clocks,
              e2e=dict(value=fps_e2e, unit="frames/s", h2d_bytes_per_step=per * 3 * SIZE * SIZE, d2h_bytes_per_step=4,
                       ms_per_step=ms_e2e / args.steps, input="uint8 frames in pinned host memory, ToTensor on the device"),
              gpu_launches=int(150 * args.steps),
              roofline=dict(bound="hbm", achieved=bytes_step / (ms_dev / args.steps * 1e-3) / 1e9, peak=pk["hbm_gbs"], unit="GB/s",
                            frac=bytes_step / (ms_dev / args.steps * 1e-3) / 1e9 / pk["hbm_gbs"], traffic=None, peak_source=pk["src"],
                            kernel="whole step",
                            note="the step is latency-bound (about 150 small launches and three host reads - proposal counts, "
                                 "labels for python's random.sample, loss - like the reference's loop); achieved = minimum bytes of "
                                 "the frozen detector forward / whole step time, a lower bound shown for scale only"),
              cpu_baseline=None))
    if world > 1:
        dist.destroy_process_group()


def run_gpu(args):
    import torch.distributed as dist
    from millieye_b200 import ops
    from millieye_b200.dist import gather_detections
    from millieye_b200.engine import capture_graph
    from oracle import darknet as odark
    from oracle.parse_config import parse_model_config
    from millieye_b200 import configs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    net, _ = build_models(device)
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    n_inputs = 3
    # Frames arrive as bytes (cv2 / camera frames, run_sp.py:205-211): the e2e arm uploads uint8 (N,3,S,S) from pinned host
    # memory and the device applies ToTensor's x / 255; the device-resident arm starts from the same values in fp32.
    host = [torch.randint(0, 256, (BATCH, 3, SIZE, SIZE), generator=gen, dtype=torch.uint8).pin_memory() for _ in range(n_inputs)]
    resident = [h.to(device).float() / 255.0 for h in host]
    plan = net.plan_for(BATCH, SIZE, device)
    from millieye_b200.models import DetectPipeline
    # N > 1: frames are independent and the path has no exchange step, so by default every rank keeps (device
    # arm) or reads back (e2e arm) the detections of its own shard and no collective runs.  ME_BENCH_GATHER=nccl gathers
    # every rank's detections on every GPU each step with ONE all_gather_into_tensor (costs 3.5 %: the NCCL kernel has to wait
    # for SMs held by the persistent conv kernels); =peer: dist.PeerGather (copy engines + stream memory
    # operations; measured slower).  ME_BENCH_HOSTALL=1 also copies a gathered batch to rank 0's host.
    mode = os.environ.get("ME_BENCH_GATHER", "0")
    gather = False if world == 1 or mode == "0" else ("peer" if mode == "peer" else True)
    pipe = DetectPipeline(net, CONF_THRESH, 0.5, 200, gather=gather,
                          host_all=rank == 0 and os.environ.get("ME_BENCH_HOSTALL", "0") == "1")
    if gather == "peer":
        try:       # CUDA IPC between the ranks' processes; fall back to the NCCL collective where it is not permitted
            pipe.submit(plan.next_input()).wait()
            ok = torch.ones(1, device=device)
        except Exception as e:  # noqa: BLE001
            print(f"[bench] rank {rank}: PeerGather unavailable ({e}); using all_gather_into_tensor", file=sys.stderr)
            ok = torch.zeros(1, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0:
            pipe = DetectPipeline(net, CONF_THRESH, 0.5, 200, gather=True, host_all=pipe.host_all)
    gather_impl = {False: "none", True: "nccl all_gather_into_tensor", "peer": "PeerGather (copy engines + stream memory ops)"}[pipe.gather]
    last = {}

    def step(x, e2e):
        # e2e arm: H2D from pinned memory on the copy stream.  Device arm: the batch already sits in the plan's input
        # buffer (DarknetPlan.next_input(), the hand-over point for on-device producers), so nothing is copied.
        # Then the graph replay; filter + NMS (+ all_gather, + D2H read of the detections in the e2e arm) follow on
        # the pipeline's second stream.
        last["rec"] = pipe.submit(x if e2e else plan.next_input(), readback=e2e)

    launches_per_step = plan_launches = None

    def timed(e2e):
        nonlocal launches_per_step, plan_launches
        src = host if e2e else resident
        if not e2e:   # inputs resident in HBM before the timed region starts: fill both input slots
            for i in range(2):
                plan.next_input().copy_(resident[i])
                pipe.submit(plan.next_input())
            torch.cuda.synchronize()
        for i in range(max(args.warmup, 3)):
            step(src[i % n_inputs], e2e)
        torch.cuda.synchronize()
        plan_launches = plan.launches
        launches_per_step = plan.launches + 2   # + nms prepare/select
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(src[i % n_inputs], e2e)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev = timed(False)
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(True)
    rec = last["rec"].wait()
    counts = rec.host_cnt.tolist()

    # conv-only time: the same launch list without decode / NMS, for the tensor roofline.  Two measurements:
    #  burst     - a few replays (tens of ms, boost clock): compared with the burst peak (a kernel timed alone)
    #  sustained - replays for >= 2 s with their own clock record: compared with the sustained peak (cuBLAS looped for
    #              seconds under the power cap), so the denominator matches how the numerator was taken
    conv_ms = conv_ms_long = None
    clocks_long = None
    if rank == 0:
        conv_ops = {i for i, b in enumerate(_op_kinds(plan)) if b == "conv"}
        torch.cuda.synchronize()
        g = capture_graph(lambda: plan.enqueue_split(only=conv_ops))
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = max(5, min(args.steps, 20))
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        conv_ms = e0.elapsed_time(e1) / reps
        n_conv = len(conv_ops) * plan.splits
        if not args.no_sustained:
            reps_long = max(reps, int(2200.0 / conv_ms) + 1)
            s2 = ClockSampler(local)
            s2.start()
            e0.record()
            for _ in range(reps_long):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            clocks_long = s2.stop()
            clocks_long["seconds"] = e0.elapsed_time(e1) * 1e-3
            conv_ms_long = e0.elapsed_time(e1) / reps_long

    # parity of what was timed: the detections of the last e2e batch's first frames against the CPU oracle
    parity = None
    if rank == 0 and not args.no_parity:
        parity = check_parity(host[(args.steps - 1) % n_inputs], rec, 4)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    flops_frame = odark.conv_flops(parse_model_config(configs.cfg_path(CFG)), SIZE)
    fps = world * BATCH * args.steps / ms_dev * 1e3
    fps_e2e = world * BATCH * args.steps / ms_e2e * 1e3
    achieved_burst = flops_frame * BATCH / (conv_ms * 1e-3) / 1e12
    achieved_long = flops_frame * BATCH / (conv_ms_long * 1e-3) / 1e12 if conv_ms_long else None
    h2d = BATCH * 3 * SIZE * SIZE     # uint8 frames
    d2h = rec.host_det.numel() * 4 + rec.host_cnt.numel() * 4      # this rank's shard (every rank reads its own)
    if rec.host_all is not None:                                      # + the gathered batch, on rank 0 only
        d2h += rec.host_all[0].numel() * 4 + rec.host_all[1].numel() * 4

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        use_ref = reference_available() and not os.environ.get("ME_BENCH_PORT")
        cb = BATCH if use_ref else CPU_BATCH
        cstep = reference_step_fn(threads, cb) if use_ref else cpu_forward_fn(threads)
        cstep()
        t0 = time.perf_counter()
        it = 0
        while it < 2 or (time.perf_counter() - t0 < 12.0 and it < 50):
            cstep()
            it += 1
        dt = time.perf_counter() - t0
        cpu = dict(value=cb * it / dt, unit="frames/s", cores=threads, kind="reference" if use_ref else "port",
                   sample=f"{it} x Darknet-53 forward+decode+NMS at batch {cb}, fp32, "
                          + ("the unmodified reference (baseline/_ref)" if use_ref else "oracle port")
                          + f", torch CPU ops, {threads} threads, {dt:.1f} s")

    # Headline fraction: the >= 2 s replay against the sustained peak when it was taken (same regime on both sides),
    # else the short replay against the burst peak; both pairs are always reported.
    if achieved_long is not None:
        achieved, peak, basis = achieved_long, pk["sustained"], "sustained (>= 2 s replay vs sustained peak)"
    else:
        achieved, peak, basis = achieved_burst, pk["burst"], "burst (short replay vs burst peak)"
    roofline = dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, basis=basis,
                    achieved_burst=achieved_burst, peak_burst=pk["burst"], frac_burst=achieved_burst / pk["burst"],
                    achieved_sustained=achieved_long, peak_sustained=pk["sustained"],
                    frac_sustained=(achieved_long / pk["sustained"]) if achieved_long else None,
                    clocks_sustained=clocks_long,
                    traffic=(measured_traffic() or {}).get("dram_bytes_per_step"), traffic_detail=measured_traffic(),
                    peak_source=pk["src"], l2_feed=l2_feed(),
                    kernel="conv_chain_kernel / conv_gemm_kernel / conv_gemm_pair_kernel / conv_first_tc_kernel (tcgen05 implicit GEMM)",
                    launches=n_conv, conv_ms_per_step=conv_ms, conv_ms_per_step_sustained=conv_ms_long,
                    note="achieved = 65.864 GFLOP/frame x 32 frames / device time of the step's conv launches (all 75 conv "
                         "layers: per-layer tcgen05 GEMMs, the tensor-core first conv and the persistent multi-layer chain "
                         "kernels), CUDA events around replays of a graph holding only those launches; traffic = DRAM bytes "
                         "(read + write) of the conv launches per step from the committed ncu capture; l2_feed = the 52-layer chain "
                         "kernel's L2->SM operand bytes per clock against the L2 slices' cap (ncu): the tensor pipe waits for "
                         "operands, the kernel sits at 94% of that cap")

    line = dict(metric=METRIC, value=fps, unit="frames/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f16",
                data="synthetic",
                config=dict(workload=f"Darknet-53 YOLOv3 inference (forward + decode + conf filter + NMS), batch {BATCH} per GPU, "
                                     f"{SIZE}x{SIZE}, fp16 compute / fp32 accumulate",
                            conf_thresh=CONF_THRESH, parallelism=f"frames sharded over {world} GPU(s), weights replicated"
                            + ("" if world == 1 else (", per-rank shards, no data-path collective" if not pipe.gather else
                                                         f", per-rank shards; detections gathered on every GPU each step: {gather_impl}")),
                            l2="2 alternating resident input batches (64 MB each; 3 rotating pinned host batches in the e2e arm) and "
                               "~4 GB of activations per step exceed the 126 MB L2; no explicit flush",
                            pipeline="DetectPipeline: filter+NMS (+all_gather, +D2H read in the e2e arm) of batch i run on a "
                                     "second stream while the convolutions of batch i+1 run; all of it inside the timed region",
                            sub_batches=plan.splits, detections_last_step=int(sum(counts))),
                clocks=clocks,
                e2e=dict(value=fps_e2e, unit="frames/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         ms_per_step=ms_e2e / args.steps,
                         input="uint8 (N,3,S,S) frames in pinned host memory, ToTensor (x / 255) on the device"),
                gpu_launches=int(launches_per_step * args.steps),
                roofline=roofline, parity_checked=bool(parity and parity.get("checked")), parity=parity,
                cpu_baseline=cpu)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def _op_kinds(plan):
    """Kind of every enqueued op, in order (recorded by DarknetPlan._add)."""
    return list(plan.op_kinds)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="darknet53", choices=["darknet53", "fusion", "train3"],
                    help="darknet53: the headline bench (BASELINE config 2); fusion: config 3; train3: config 4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s conv-only replay")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed detections")
    ap.add_argument("--obj-bias", type=float, default=None, help="synthetic head statistics: objectness bias (experiments)")
    ap.add_argument("--head-gain", type=float, default=None, help="synthetic head statistics: head weight gain (experiments)")
    args = ap.parse_args()
    if args.obj_bias is not None:
        WEIGHTS["obj_bias"] = args.obj_bias
    if args.head_gain is not None:
        WEIGHTS["head_gain"] = args.head_gain
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (use --impl reference for the CPU arm)")
        {"darknet53": run_gpu, "fusion": run_fusion, "train3": run_train3}[args.config](args)


if __name__ == "__main__":
    main()
