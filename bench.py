#!/usr/bin/env python
"""Benchmark of the FusionDepth per-step training hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on host cores

Metric (BASELINE.json): training images/sec at 640x192, batch 12 (the reference runs
`--batch_size 12` as two micro-batches of 6 with loss/2 and one Adam step, trainer.py:30-41,
237-248).  One "step" = one optimiser step = 12 images per GPU: 6 ResNet-18 trunks +
DepthDecoder + 2 PoseDecoders forward/backward per micro-batch, the fused photometric loss
chain, gradient all-reduce (N>1) and Adam.  Prints ONE JSON line (rank 0).
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

# The captured step has up to 14 concurrent branches (2 micro-batches x (1 + 5 trunk streams) + copies);
# the default 8 hardware work queues alias them (measured +3 % with 32, the maximum).  Must be set before
# the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402
import torch  # noqa: E402

# --workload r18 (default) is BASELINE.json's metric / configs[1]; the other two are its configs[2] and
# configs[4], selectable so that they have a measured line too (profiles/), not part of the driver's run.
WORKLOADS = {
    "r18": dict(H=192, W=640, MICRO_B=6, ACCUM=2, NUM_LAYERS=18, KIND="trainer",
                METRIC="training images/sec (640x192 b12)",
                WORKLOAD="trainer.py step (enc/dec/pose fwd+bwd + reprojection loss + Adam), "
                         "ResNet-18, 640x192, batch 12 per GPU = 2 micro-batches of 6"),
    "r50": dict(H=320, W=1024, MICRO_B=8, ACCUM=1, NUM_LAYERS=50, KIND="trainer",
                METRIC="training images/sec (1024x320 b8, ResNet-50)",
                WORKLOAD="trainer.py step (enc/dec/pose fwd+bwd + reprojection loss + Adam), "
                         "ResNet-50, 1024x320, batch 8 per GPU = 1 micro-batch of 8 (trainer.py:30-41)"),
    "refiner": dict(H=192, W=640, MICRO_B=6, ACCUM=1, NUM_LAYERS=18, KIND="refiner",
                    METRIC="refiner images/sec (640x192, --batch_size 12 = one optimiser step per 6 images)",
                    WORKLOAD="refiner.py step (frozen stage-1 nets fwd, pseudo-3D pack, pose nets, refine2d "
                             "decoder fwd+bwd, reprojection + GDC si-loss, Adam), ResNet-18, 640x192, "
                             "6 images per optimiser step (refiner.py:34-45, 270-278)"),
}
H, W, MICRO_B, ACCUM, NUM_LAYERS, KIND = 192, 640, 6, 2, 18, "trainer"
METRIC, WORKLOAD = WORKLOADS["r18"]["METRIC"], WORKLOADS["r18"]["WORKLOAD"]


def select_workload(name: str):
    globals().update(WORKLOADS[name])


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "tf": p["bf16_tflops_sustained"], "tf_burst": p.get("bf16_tflops", p["bf16_tflops_sustained"]),
                "src": "measured"}
    return {"hbm": 6650.0, "tf": 1400.0, "tf_burst": 1650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_step_inputs(seed: int, device=None):
    """ACCUM micro-batches of MICRO_B images (CPU tensors); 4-beam scans go through the CUDA LiDAR
    kernels when a device is given, else through the stand-in."""
    from fusiondepth_b200 import synth
    lidar_fn = None
    if device is not None and (H, W) == (192, 640):       # the reference's LiDAR data layer is 192x640-only
        from fusiondepth_b200 import lidar
        P = synth.velo_to_image_matrix(synth.parse_roundtrip(), 2)

        def lidar_fn(points):
            fb, two = lidar.lidar_inputs([points], [P], device=device)
            return {"4beam": fb[0].cpu(), "2channel": two[0].cpu()}
    batches, noises = [], []
    for i in range(ACCUM):
        # ranges around 26 x (random-init depth ~0.2) = 5.2 m keep the si-loss mask populated, which
        # the reference needs for a finite loss (trainer.py:584-587); KITTI ranges would empty it
        make = synth.make_refiner_batch if KIND == "refiner" else synth.make_batch
        b = make(MICRO_B, H, W, seed=seed * 10 + i, lidar_fn=lidar_fn, mode="coherent",
                 lidar_density=0.03, scan_range=(3.0, 10.0))
        noises.append(b.pop("noise"))
        batches.append(b)
    return batches, noises


# ------------------------------------------------------------------------------------------------
def reference_stepper(device, allow_tf32=None):
    """The UNMODIFIED reference (oracle/_ref staged by oracle/make_ref.py, or /root/reference) wired for one
    micro-batch: returns (step_fn(batch, noise) -> loss tensor, kind).  step_fn = zero_grad, the reference's own
    Trainer/Refiner.process_batch, (loss / accumulate).backward(), torch.optim.Adam.step() -- trainer.py:237-248,
    refiner.py:270-278.  Falls back to the oracle port (kind "port") when the sources are not staged."""
    from oracle import ref_harness as RH
    from tests._util import synth_weights
    lr = 1e-4 * (MICRO_B * ACCUM) / 8
    if allow_tf32 is not None:
        torch.backends.cudnn.allow_tf32 = bool(allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
    if RH.available():
        ns = RH.load(with_refiner=(KIND == "refiner"))
        torch.manual_seed(0)
        models = RH.make_models(ns, NUM_LAYERS)
        if KIND == "refiner":
            models["refine2d_decoder"] = RH.make_refine_decoder(ns, models["encoder"].num_ch_enc)
        for m in models.values():
            m.to(device).train()
        if KIND == "refiner":
            drv = RH.make_refiner(ns, models, MICRO_B, H, W, device=device)
            params = list(models["refine2d_decoder"].parameters())
        else:
            drv = RH.make_trainer(ns, models, MICRO_B, H, W, device=device, batch_size_flag=MICRO_B * ACCUM)
            params = [p for m in models.values() for p in m.parameters()]
        optim = torch.optim.Adam(params, lr)

        def step(batch, noise):
            optim.zero_grad()
            with RH.FixedNoise([noise[s] for s in range(4)], cpu=(str(device) == "cpu")):
                _, losses = drv.process_batch(dict(batch))
            (losses["loss"] / ACCUM).backward()
            optim.step()
            return losses["loss"].detach()
        return step, "reference"
    if str(device) != "cpu":
        raise RuntimeError("reference sources not staged (oracle/_ref); the oracle port is CPU-only")
    from oracle import step_oracle as SO
    from fusiondepth_b200 import refine, training
    torch.manual_seed(0)
    models = (refine.build_refiner_models if KIND == "refiner" else training.build_models)(NUM_LAYERS, "cpu")
    sds = {}
    for name, m in models.items():
        train = KIND != "refiner" or name == "refine2d_decoder"
        sds[name] = {k: (v.detach().clone().contiguous().to(device).requires_grad_(True)
                         if train and v.is_floating_point() and "running" not in k else v.detach().clone().to(device))
                     for k, v in m.state_dict().items()}
    params = [t for sd in sds.values() for t in sd.values() if t.requires_grad]
    m1, m2 = [torch.zeros_like(p) for p in params], [torch.zeros_like(p) for p in params]
    it = [0]

    def step(batch, noise):
        for p in params:
            p.grad = None
        fn = SO.refiner_process_batch if KIND == "refiner" else SO.process_batch
        _, losses = fn(sds, batch, noise, NUM_LAYERS, True)
        (losses["loss"] / ACCUM).backward()
        it[0] += 1
        with torch.no_grad():
            live = [(p, a, b) for p, a, b in zip(params, m1, m2) if p.grad is not None]
            SO.adam_step([x[0] for x in live], [x[0].grad for x in live], [x[1] for x in live],
                         [x[2] for x in live], it[0], lr)
        return losses["loss"].detach()
    return step, "port"


def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores: the UNMODIFIED
    Trainer.process_batch (staged sources, see oracle/make_ref.py) + backward + Adam, all host threads.
    Each step = one micro-batch of MICRO_B images (a bounded sample of the workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    step, kind = reference_stepper("cpu")
    batches, noises = synthetic_step_inputs(1)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        step(batches[it % ACCUM], noises[it % ACCUM])
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = MICRO_B / (ms / 1e3)
    sample = ("1 micro-batch of %d images (fwd+bwd+Adam) per step, %s, fp32"
              % (MICRO_B, "unmodified reference process_batch" if kind == "reference" else "oracle port"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "cuda_graph": False, "parallelism": "host cpu, %d threads" % cores,
                   "l2": "n/a (host)"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_pytorch_gpu(args):
    """Informative arm: the UNMODIFIED reference on the same B200 through
    stock PyTorch (cuDNN / ATen kernels, eager, no graph), fp32 with cudnn.allow_tf32 False and True.  Inputs
    resident in HBM; one step = one optimiser step of the workload (ACCUM micro-batches + Adam)."""
    from fusiondepth_b200 import synth
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    out = {"impl": "pytorch-gpu", "metric": METRIC, "unit": "images/s", "config": {"workload": WORKLOAD},
           "steps": args.steps, "warmup": max(args.warmup, 3), "modes": {}}
    cpu_batches, cpu_noises = synthetic_step_inputs(100)
    batches = [synth.to_device(b, dev) for b in cpu_batches]
    for name, tf32 in (("fp32", False), ("tf32", True)):
        try:
            step, kind = reference_stepper(dev, allow_tf32=tf32)
        except RuntimeError as ex:
            print(json.dumps({"impl": "pytorch-gpu", "unavailable": str(ex)}))
            return
        out["kind"] = kind

        def opt_step():
            for b, n in zip(batches, cpu_noises):          # the reference draws its noise on the CPU too
                loss = step(b, n)
            return loss
        for _ in range(max(args.warmup, 3)):
            opt_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = opt_step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out["modes"][name] = {"value": MICRO_B * ACCUM / (ms * 1e-3), "ms_per_step": ms, "loss": float(loss),
                              "cudnn_allow_tf32": tf32}
        del step
        torch.cuda.empty_cache()
    out["value"] = out["modes"]["fp32"]["value"]
    print(json.dumps(out))


def ncu_traffic(which="conv"):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r02_ncu_traffic.json, from the raw csv export next to it); None if absent."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    try:
        return json.load(open(path)).get(which)
    except Exception:
        return None


def dominant_kernel_leg(dev):
    """The most frequent tensor-core conv of the step (ResNet layer1 3x3 64->64 on 6x48x160, 112 of
    the ~390 forward / data-gradient conv launches per step are this shape): average launch duration
    over 3 x 20 back-to-back launches captured in a CUDA graph, CUDA events, inputs L2-warm.  The kernel is
    launched through the C-ABI entry point the step uses (fd_conv2d_fwd_tc) with the low-order weight tensor
    prepared once, as the step's weight cache does -- ops.conv2d without that cache would add one fd_tf32_split
    launch per call to the timed region."""
    import ctypes
    from fusiondepth_b200 import _lib
    lib = _lib.load()
    B, C, H, W = MICRO_B, 64, globals()["H"] // 4, globals()["W"] // 4
    x = torch.randn(B, H, W, C, device=dev)                       # NHWC
    w = torch.randn(C, 3, 3, C, device=dev)                       # [Cout, KH, KW, Cin]
    wlo = torch.empty_like(w)
    y = torch.empty(B, H, W, C, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())

    def launch():
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.fd_conv2d_fwd_tc(P(x), P(w), P(wlo), None, P(y), B, H, W, C, C, 3, 3, 1, 1, 0, st),
                   "fd_conv2d_fwd_tc")

    with torch.cuda.stream(s):
        _lib.check(lib.fd_tf32_split(P(w), P(wlo), w.numel(), ctypes.c_void_p(s.cuda_stream)), "fd_tf32_split")
        launch()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            launch()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 60
    flop = 2.0 * B * H * W * C * C * 9
    return {"kernel": "conv_tc4_kernel<64,0> (persistent tcgen05 implicit GEMM; layer1 3x3 64->64, M=%d, K=576)"
                      % (B * H * W), "avg_us": us, "flop_per_launch": flop,
            "achieved_tflops": flop / (us * 1e-6) / 1e12, "mma_tflops_issued": 3 * flop / (us * 1e-6) / 1e12}


def lidar_leg(dev, n_frames=36):
    """The sparse-LiDAR input kernels (kitti_utils.generate_depth_map -> max_pool2d/100 -> get_4beam_2channel)
    for the 36 scans of one 12-image step (3 frames each), timed alone with CUDA events.  They run once per
    batch in input preparation (the reference caches their result as .npy), so they are NOT inside the timed
    step; algorithmic bytes per SURVEY.md section 8(d): 16 B/point + the 384x1280 fp64 map; 192x640x4 B in,
    2x that out."""
    from fusiondepth_b200 import lidar, synth
    P = synth.velo_to_image_matrix(synth.parse_roundtrip(), 2)
    scans = [synth.make_scan(1000 + i) for i in range(n_frames)]
    Ps = [P] * n_frames
    lidar.lidar_inputs(scans, Ps, device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        lidar.lidar_inputs(scans, Ps, device=dev)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    npts = sum(s.shape[0] for s in scans)
    bytes_alg = 16.0 * npts + n_frames * (384 * 1280 * 8 + 192 * 640 * 4 * 3)
    pk = peaks()
    return {"kernel": "fd_lidar_depth_map + fd_lidar_pool_scale + fd_two_channel, %d scans (%d points), incl. the "
                      "host->device copy of the points" % (n_frames, npts),
            "bound": "hbm", "us_per_batch": us, "achieved": bytes_alg / (us * 1e-6) / 1e9, "peak": pk["hbm"],
            "unit": "GB/s", "frac": bytes_alg / (us * 1e-6) / 1e9 / pk["hbm"], "traffic": None,
            "in_timed_step": False}


def eval_leg(models, dev, iters=50):
    """Inference path (evaluate_depth.py:173-237 wiring): batch-1 latency of encoder + beam
    encoder + depth decoder with BatchNorm folded into the convolutions, one CUDA-graph replay per frame, input
    copied in and disparity produced on the device.  Informative; not part of the timed training step."""
    from fusiondepth_b200 import evaluation
    keep = {k: m.training for k, m in models.items()}
    try:
        run = evaluation.EvalRunner({k: models[k] for k in ("encoder", "beam_encoder", "depth")}, 1, H, W)
        x = {("color", 0, 0): torch.rand(1, 3, H, W, device=dev), "2channel": torch.rand(1, 2, H, W, device=dev)}
        for _ in range(5):
            run(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            run(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        return {"what": "batch-1 inference (encoder + beam encoder + depth decoder, BN folded, CUDA graph), %dx%d" % (W, H),
                "latency_ms": ms, "frames_per_s": 1e3 / ms}
    finally:
        for k, m in models.items():
            m.train(keep[k])


def cpu_baseline_leg(sds0, batches, noises):
    """Bounded CPU sample on rank 0: one optimiser step (ACCUM micro-batches, fwd+bwd) after one warm-up
    micro-batch -- on the SAME initial weights and the SAME batch as the CUDA arm's first step, so its loss is
    the parity reference for `loss_first`.  Runs the UNMODIFIED reference's process_batch when its sources are
    staged (kind "reference"), else the oracle port (kind "port")."""
    from oracle import ref_harness as RH
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    if RH.available():
        kind = "reference"
        ns = RH.load(with_refiner=(KIND == "refiner"))
        models = RH.make_models(ns, NUM_LAYERS)
        if KIND == "refiner":
            models["refine2d_decoder"] = RH.make_refine_decoder(ns, models["encoder"].num_ch_enc)
        for name, m in models.items():
            m.load_state_dict(sds0[name])
            m.train()
        drv = (RH.make_refiner(ns, models, MICRO_B, H, W) if KIND == "refiner" else
               RH.make_trainer(ns, models, MICRO_B, H, W, batch_size_flag=MICRO_B * ACCUM))

        def micro(b, n):
            with RH.FixedNoise([n[s] for s in range(4)], cpu=True):
                _, l = drv.process_batch(dict(b))
            (l["loss"] / ACCUM).backward()
            return float(l["loss"])
    else:
        kind = "port"
        from oracle import step_oracle as SO
        train = lambda name: KIND != "refiner" or name == "refine2d_decoder"
        sds = {name: {k: (v.detach().clone().contiguous().requires_grad_(True)
                          if train(name) and v.is_floating_point() and "running" not in k else v.detach().clone())
                      for k, v in sd.items()} for name, sd in sds0.items()}
        fn = SO.refiner_process_batch if KIND == "refiner" else SO.process_batch

        def micro(b, n):
            _, l = fn(sds, b, n, NUM_LAYERS, True)
            (l["loss"] / ACCUM).backward()
            return float(l["loss"])
    micro(batches[0], noises[0])                       # warm-up (does not change the weights)
    t0 = time.perf_counter()
    total = 0.0
    for b, n in zip(batches, noises):
        total += micro(b, n) / ACCUM
    dt = time.perf_counter() - t0
    return {"value": MICRO_B * ACCUM / dt, "unit": "images/s", "cores": cores, "kind": kind,
            "sample": "1 optimiser step (%d micro-batch(es) of %d, fwd+bwd, %s) after 1 warm-up micro-batch"
                      % (ACCUM, MICRO_B, "unmodified reference process_batch" if kind == "reference" else "oracle port"),
            "oracle_loss_first_step": total}


def sub_bench(extra_args, env=None, timeout=600):
    """Runs another arm of this script in a fresh process (its own CUDA context and libraries) and returns
    its JSON line, or {"unavailable": why}."""
    cmd = [sys.executable, os.path.abspath(__file__)] + extra_args
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        e.pop(k, None)
    e.update(env or {})
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e)
        for line in reversed(r.stdout.strip().split("\n")):
            if line.startswith("{"):
                return json.loads(line)
        return {"unavailable": "rc=%d %s" % (r.returncode, r.stderr.strip()[-300:])}
    except Exception as ex:                       # noqa: BLE001
        return {"unavailable": repr(ex)[:300]}


# ------------------------------------------------------------------------------------------------
def _phase(msg):
    if os.environ.get("FD_BENCH_VERBOSE"):
        sys.stderr.write("[bench rank %s] %s\n" % (os.environ.get("RANK", "0"), msg))
        sys.stderr.flush()


def run_ours(args):
    from fusiondepth_b200 import _lib, ops, refine, synth, training
    _lib.load()                                   # no CUDA extension => fail loudly
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)

    _phase("process group up")
    torch.manual_seed(0)                          # same initial weights on every rank
    if KIND == "refiner":
        models = refine.build_refiner_models(NUM_LAYERS, dev)
    else:
        models = training.build_models(NUM_LAYERS, dev)
    sds0 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sds0 = {name: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
                for name, m in models.items()}
    lr = 1e-4 * (MICRO_B * ACCUM) / 8
    if KIND == "refiner":
        step = refine.RefineStep(models, lr=lr)
    else:
        step = training.TrainStep(models, lr=lr, accumulate=ACCUM)
    # Every rank feeds its own copy of the same seeded batch: the content does not affect the timing, and with
    # random-initialised networks distinct random batches empty the si-loss mask on some rank after the first
    # Adam step (loss = NaN there -- in the reference too), which the gradient all-reduce
    # would spread to every replica.
    cpu_batches, cpu_noises = synthetic_step_inputs(100, dev)
    pinned = [{k: v.pin_memory() for k, v in b.items()} for b in cpu_batches]
    pinned_noise = [{k: v.pin_memory() for k, v in n.items()} for n in cpu_noises]
    h2d = sum(v.numel() * v.element_size() for b in pinned for v in b.values()) + \
        sum(v.numel() * v.element_size() for n in pinned_noise for v in n.values())
    batches = [synth.to_device(b, dev) for b in cpu_batches]
    noises = [{s: t.to(dev) for s, t in n.items()} for n in cpu_noises]

    _phase("inputs ready")

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    if args.profile_step:
        step.step(batches, noises)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step.step(batches, noises)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return

    n0 = _lib.launch_count()
    first_loss = None
    if args.no_graph:
        run = lambda: step.step(batches, noises)
        first_loss = float(run())
        launches_per_step = _lib.launch_count() - n0
    else:
        step.capture(batches, noises, warmup=1)
        launches_per_step = (_lib.launch_count() - n0) // 2      # 1 warm-up + 1 capture pass
        run = step.replay

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            out = fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t) / k, out

    _phase("captured / first step done, launches/step=%d" % launches_per_step)
    if first_loss is None:
        first_loss = float(run())               # the captured step starts from the initial weights
    else:
        run()
    _phase("first replay done")
    for _ in range(max(args.warmup, 3) - 1):
        run()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, loss = timed(run, args.steps)
    _phase("timed region done: %.2f ms/step" % ms)
    clocks = sampler.stop() if rank == 0 else None

    # End to end: every step copies ITS inputs from pinned host memory (H2D) and reads its loss back (D2H)
    # inside the timed region.  With the captured graph the copy of step k+1 is issued on a copy stream
    # while step k computes (TrainStep.feeder(): a prefetching loader); the first step's copy is not
    # hidden and K copies are made for K steps.
    host_loss = torch.empty(1).pin_memory()

    def e2e_eager():
        b = [synth.to_device(x, dev) for x in pinned]
        n = [{s: t.to(dev, non_blocking=True) for s, t in x.items()} for x in pinned_noise]
        out = step.step(b, n)
        host_loss.copy_(out.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_loss

    if args.no_graph:
        e2e_eager()
        ms_e2e, _ = timed(e2e_eager, args.steps)
    else:
        feeder = step.feeder()

        host_losses = [torch.empty(1).pin_memory() for _ in range(2)]
        loss_on_host = [torch.cuda.Event() for _ in range(2)]

        def e2e_run(k):
            """Every step: H2D of ITS inputs (prefetched under the previous step), graph replay, D2H of ITS loss.
            The host waits for the loss of step i-1 after it has enqueued step i -- what a training loop that logs
            the previous step's loss does -- so the launch of an 1800-node graph is not exposed every step."""
            feeder.prefetch(pinned, pinned_noise)
            for i in range(k):
                feeder.commit()
                if i + 1 < k:
                    feeder.prefetch(pinned, pinned_noise)
                out = step.replay()
                host_losses[i & 1].copy_(out.reshape(1), non_blocking=True)
                loss_on_host[i & 1].record()
                if i > 0:
                    loss_on_host[(i - 1) & 1].synchronize()
            loss_on_host[(k - 1) & 1].synchronize()
            host_loss.copy_(host_losses[(k - 1) & 1])
            return host_loss

        e2e_run(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(args.steps)
        e1.record()
        barrier()
        t_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t_e2e, op=torch.distributed.ReduceOp.MAX)
        ms_e2e = float(t_e2e) / args.steps
    _phase("e2e done")
    replica_drift = None
    if world > 1:
        # every rank applied the same averaged gradients to the same initial weights: the replicas must agree
        ref = step.flat.data.clone()
        torch.distributed.broadcast(ref, 0)
        d = (ref - step.flat.data).abs().max().reshape(1)
        torch.distributed.all_reduce(d, op=torch.distributed.ReduceOp.MAX)
        replica_drift = float(d)

    # Per-kernel-family device times: ONE instrumented eager step on rank 0, on a single stream (no
    # trunk / micro-batch concurrency, no collective) with CUDA events around every conv and loss
    # launch, so each event pair brackets one kernel running alone; the whole serial step is timed too.
    fam, serial_ms, dom, lidar_rf, eval_rf = {}, None, None, None, None
    if rank == 0:
        ops.PROFILE = {}
        step.world = 1
        keep = (step.trunks, step.concurrent)
        step.trunks, step.concurrent = [None], False
        step.step(batches, noises)                        # warm (allocator) on this code path
        torch.cuda.synchronize()
        ops.PROFILE = {}
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        step.step(batches, noises)
        s1.record()
        torch.cuda.synchronize()
        serial_ms = s0.elapsed_time(s1)
        for name, evs in ops.PROFILE.items():
            fam[name] = (sum(s.elapsed_time(e) for s, e, _ in evs), sum(w for _, _, w in evs), len(evs))
        ops.PROFILE = None
        step.trunks, step.concurrent = keep
        dom = dominant_kernel_leg(dev)
        lidar_rf = lidar_leg(dev)
        try:
            eval_rf = eval_leg(models, dev) if KIND == "trainer" else None
        except Exception as ex:                     # noqa: BLE001  (informative leg: never lose the bench line)
            eval_rf = {"unavailable": repr(ex)[:200]}
    _phase("instrumented step done")

    result = None
    if rank == 0:
        imgs = MICRO_B * ACCUM * world
        pk = peaks()
        conv_ms, conv_flop, conv_n = fam.get("conv", (0.0, 0.0, 0))
        pl_ms = fam["photoloss_fwd"][0] + fam["photoloss_bwd"][0]
        pl_bytes = fam["photoloss_fwd"][1] + fam["photoloss_bwd"][1]
        conv_tf = conv_flop / (conv_ms * 1e-3) / 1e12 if conv_ms else 0.0
        pl_gbs = pl_bytes / (pl_ms * 1e-3) / 1e9
        traffic = ncu_traffic()
        # the dominant kernel, timed alone (burst peak); the whole conv family inside the step next to it (sustained)
        roofline = {"kernel": dom["kernel"], "bound": "tensor", "achieved": dom["achieved_tflops"],
                    "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": dom["achieved_tflops"] / pk["tf_burst"],
                    "traffic": traffic, "peak_source": pk["src"], "avg_us": dom["avg_us"],
                    "flop_per_launch": dom["flop_per_launch"], "mma_tflops_issued": dom["mma_tflops_issued"],
                    "peak_note": "peak = measured dense bf16 cuBLAS (burst: the kernel is timed alone); this path "
                                 "issues kind::tf32 MMAs (half the bf16 rate) three times per product (3xTF32 for "
                                 "fp32 parity), so its ceiling is peak/6",
                    "frac_of_3xtf32_ceiling": dom["achieved_tflops"] / (pk["tf_burst"] / 6.0),
                    "family": {"kernel": "conv implicit-GEMM family inside the step (conv_tc4 / conv_tc3 / conv_tc2 "
                                         "fwd + dgrad, conv_wgrad2 on tcgen05, small-channel CUDA-core kernels), "
                                         "%d launches/step" % conv_n,
                               "achieved": conv_tf, "peak": pk["tf"], "frac": conv_tf / pk["tf"],
                               "frac_of_3xtf32_ceiling": conv_tf / (pk["tf"] / 6.0),
                               "ms_per_step": conv_ms, "serial_step_ms": serial_ms,
                               "share_of_serial_step": conv_ms / serial_ms if serial_ms else None}}
        roofline_loss = {"kernel": "fd_photoloss_fwd+bwd (4 launches/step)", "bound": "hbm",
                         "achieved": pl_gbs, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": pl_gbs / pk["hbm"], "traffic": ncu_traffic("photoloss"),
                         "peak_source": pk["src"], "ms_per_step": pl_ms,
                         "share_of_serial_step": pl_ms / serial_ms if serial_ms else None}
        result = {
            "metric": METRIC, "value": imgs / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "cuda_graph": not args.no_graph, "parallelism": "dp%d" % world,
                       "cuda_device_max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"),
                       "l2": "no explicit flush: each step streams >2 GB of activations (>> 126 MB L2)"},
            "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "how": "pinned host batch -> H2D on a copy stream (prefetch of step k+1 under step k) -> "
                           "graph replay -> D2H loss of every step; the host waits for the loss of step k after it has "
                           "enqueued step k+1 (as a loop that logs the previous step's loss does)"},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roofline, "roofline_loss": roofline_loss, "roofline_lidar": lidar_rf,
            "inference": eval_rf,
            "loss_first": first_loss, "loss": float(loss),
            "conv_precision": os.environ.get("FD_CONV_PRECISION", "3xtf32"),
        }
        if world > 1:
            result["dp"] = {"allreduce": "bucketed, overlapped with backward" if step.bucketed else "single, after backward",
                            "buckets_floats": [hi - lo for _, lo, hi in step.buckets],
                            "max_abs_weight_difference_between_replicas": replica_drift}
        parity_ok = True
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline_leg(sds0, cpu_batches, cpu_noises)
            want = cb.pop("oracle_loss_first_step")
            rel = abs(first_loss - want) / abs(want)
            parity_ok = rel < 1e-4
            result["cpu_baseline"] = cb
            result["parity"] = {"what": "loss of the first optimiser step (12 images, initial weights): CUDA "
                                        "path vs the CPU oracle on the identical batch",
                                "loss_first": first_loss, "oracle_loss": want, "rel_err": rel, "tol": 1e-4,
                                "ok": parity_ok}
        if world == 1 and not args.no_extras and not args.no_graph:
            # informative arms, each in its own process: stock PyTorch (cuDNN) running the unmodified reference on
            # this GPU, and this repo's single-pass-TF32 fast mode (the arithmetic cuDNN uses with allow_tf32)
            _phase("extras")
            wl = ["--workload", args.workload]
            pg = sub_bench(["--impl", "pytorch-gpu", "--steps", "5", "--warmup", "3"] + wl)
            result["pytorch_gpu"] = {k: pg.get(k) for k in ("modes", "kind", "unavailable") if k in pg}
            fm = sub_bench(["--steps", str(args.steps), "--warmup", "3", "--no-cpu-baseline", "--no-extras"] + wl,
                           env={"FD_CONV_PRECISION": "tf32"})
            if "value" in fm:
                ref_loss = result.get("parity", {}).get("oracle_loss")
                result["fast_mode"] = {
                    "what": "FD_CONV_PRECISION=tf32: single-pass TF32 tensor-core convolutions (1 MMA per product "
                            "instead of 3); NOT the headline -- it does not meet the 1e-4 parity bar",
                    "value": fm["value"], "unit": fm["unit"], "ms_per_step": fm["ms_per_step"],
                    "e2e": fm["e2e"]["value"], "loss_first": fm["loss_first"],
                    "loss_rel_err_vs_fp32_oracle": (abs(fm["loss_first"] - ref_loss) / abs(ref_loss)
                                                    if ref_loss else None)}
            else:
                result["fast_mode"] = fm
    if result is not None:
        print(json.dumps(result))
        sys.stdout.flush()
        if not parity_ok:
            sys.stderr.write("bench.py: PARITY FAILURE: loss_first %.9g vs oracle %.9g (rel %.3g > 1e-4)\n"
                             % (first_loss, want, rel))
            sys.stderr.flush()
            if world == 1:
                sys.exit(3)
    if world > 1:
        # A communicator that is referenced by a captured CUDA graph does not tear down cleanly
        # (destroy_process_group blocks); the line is out, so leave without the NCCL teardown.
        torch.cuda.synchronize()
        _phase("exit")
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "pytorch-gpu"])
    ap.add_argument("--workload", default="r18", choices=sorted(WORKLOADS),
                    help="r18 = BASELINE.json's metric (default); r50 / refiner = its configs 3 and 5")
    ap.add_argument("--no-extras", dest="no_extras", action="store_true",
                    help="skip the informative pytorch-gpu and fast-mode sub-runs")
    ap.add_argument("--no-graph", dest="no_graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--profile-step", dest="profile_step", action="store_true",
                    help="run ONE eager optimiser step between cudaProfilerStart/Stop (for ncu "
                         "--profile-from-start off) and exit; prints no bench line")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "pytorch-gpu":
        run_pytorch_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
