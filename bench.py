#!/usr/bin/env python
"""Benchmark of the DeNet training hot path on B200 (BASELINE.json metric: DeNet-34 training images/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload denet34-skip]

One "step" = one ModelCNN.train_step on one synthetic batch (forward, on-device RoI sampling, backward, gradient
all-reduce when N > 1, solver update).  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the meaning
of every field.  `--impl reference` times the reference's algorithm on the host CPU cores (oracle/: the compiled
reference C++ sampler + the torch-CPU restatement of the Theano graph; Theano itself cannot be installed offline).
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "denet34_train_images_per_sec"
UNIT = "images/s"
SOLVER_HP = dict(lr=0.1, momentum=[0.9, 0.9], decay=1e-4)     # papers/dss/denet34.sh:43


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="denet34-skip")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-images", type=int, default=2, help="images per CPU-baseline step")
    ap.add_argument("--cpu-sample-steps", type=int, default=4)
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--detail", action="store_true", help="also print a per-layer conv timing table to stderr")
    ap.add_argument("--no-parity-mode", action="store_true",
                    help="skip the fp32-parity-mode measurement (a second model, bf16x3 MMA) reported beside the bf16 line")
    ap.add_argument("--parity-steps", type=int, default=5)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ synthetic data
def synthetic_batch(batch, data_shape, classes, seed):
    """images U(0,1) fp32 NCHW; 1-8 GT boxes per image, x0,y0~U(0,0.7), w,h~U(0.1,0.3)"""
    import numpy
    rs = numpy.random.RandomState(seed)
    x = rs.uniform(0, 1, (batch,) + tuple(data_shape)).astype(numpy.float32)
    rnd = random.Random(seed)
    metas = []
    for _ in range(batch):
        boxes, cls = [], []
        for _ in range(rnd.randint(1, 8)):
            x0, y0 = rnd.uniform(0, 0.7), rnd.uniform(0, 0.7)
            boxes.append((x0, y0, min(1.0, x0 + rnd.uniform(0.1, 0.3)), min(1.0, y0 + rnd.uniform(0.1, 0.3))))
            cls.append(rnd.randint(0, classes - 1))
        metas.append({"bbox": boxes, "class": cls, "image_class": cls[0]})
    return x, metas


def build_model(workload, batch, seed=1):
    import numpy
    from denet_b200.model import model_cnn, recipes
    desc, data_shape, wb, classes, convert, solver = recipes.WORKLOADS[workload]
    batch = batch or wb
    numpy.random.seed(seed)
    model = model_cnn.ModelCNN()
    model.batch_size, model.class_num = batch, classes
    model.build(desc.split(), data_shape, "relu", "half", ["he-backward"])
    if convert:
        model.convert_bn_relu()
    return model, data_shape, batch, classes, solver


def conv_flops(model):
    """algorithmic conv FLOPs of one training step (fprop + dgrad + wgrad, no dgrad for the first conv)"""
    from denet_b200.model.model_cnn import _walk
    convs = [l for l in _walk(model.layers) if l.type_name == "conv" and l.enabled]
    fwd = sum(l.fprop_flops() for l in convs)
    nodgrad = sum(l.fprop_flops() for l in convs if l.is_first)
    return fwd, 3 * fwd - nodgrad, len(convs)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_run(workload, images_per_step, steps, warmup, seed=1):
    """the reference's training iteration on the host cores through oracle.ref_train.RefTrainer -> images/s"""
    import torch
    from oracle.ref_train import RefTrainer
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, data_shape, _, classes, solver = build_model(workload, images_per_step, seed)
    js = model.export_json()["layers"]          # host-side description + initial weights only; no kernels involved
    del model
    random.seed(seed)
    trainer = RefTrainer(js, (images_per_step,) + tuple(data_shape), classes, solver=solver, dtype=torch.float32)
    x, metas = synthetic_batch(images_per_step, data_shape, classes, seed)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        cost, _ = trainer.train_step(x, metas, it, SOLVER_HP["lr"], SOLVER_HP["momentum"], SOLVER_HP["decay"])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": images_per_step * len(times) / total, "unit": UNIT, "cores": cores,
            "kind": "port", "ms_per_step": 1000.0 * total / len(times),
            "sample": "%d step(s) of %d image(s) of the %s workload (two forward passes + backward + update per "
                      "step as in the reference; torch-CPU fp32 restatement of the Theano graph, RoI sampling by %s)"
                      % (len(times), images_per_step, workload, trainer.sample_impl)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    images = args.cpu_sample_images
    from denet_b200.model import recipes
    _, data_shape, wb, classes, _, solver = recipes.WORKLOADS[args.workload]
    res = cpu_reference_run(args.workload, images, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the CPU arm runs a bounded SAMPLE of the workload: `images` images per step, stated as such (the full
            # per-GPU batch of the GPU arm would take minutes per step on the host cores)
            "config": dict(workload_config(args.workload, images, 1, data_shape, classes, solver),
                           sample_of="per-GPU batch %d of the GPU arm" % (args.batch or wb)),
            "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                             "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(workload, batch, gpus, data_shape, classes, solver):
    return {"workload": "%s training step, synthetic %dx%dx%d images, %d classes, per-GPU batch %d"
                        % (workload, data_shape[0], data_shape[1], data_shape[2], classes, batch),
            "per_gpu_batch": batch, "global_batch": batch * gpus, "solver": solver,
            "parallelism": "dp%d (batch sharded over ranks, NCCL gradient all-reduce)" % gpus if gpus > 1 else "single GPU",
            "l2": "working set per step (activations + gradients, several GB) far exceeds the 126 MB L2; no flush needed"}


def measure_parity_mode(args, dev, x_dev, metas, peak):
    """images/s, ms/step and conv TFLOP/s of the fp32-parity mode (bf16x3 MMA: a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with
    fp32 accumulation).  Its conv FLOP figure counts the ALGORITHMIC FLOPs once (the tensor pipe executes 3x as many)."""
    import torch
    from denet_b200 import lib
    model, data_shape, batch, classes, solver = build_model(args.workload, args.batch, seed=1)
    model.to_device(dev, precision="fp32")
    model.build_train_func(solver, [])
    random.seed(1)
    hp = SOLVER_HP
    fwd_flops, train_flops, _ = conv_flops(model)
    it = 0

    def step():
        nonlocal it
        model._train_step_device(x_dev, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"])
        it += 1
    names = ["denet_conv2d_fprop", "denet_conv2d_fprop_scatter", "denet_conv2d_dgrad_bnbwd", "denet_conv2d_wgrad",
             "denet_conv2d_rowfold_fprop", "denet_conv2d_rowfold_wgrad", "denet_wgrad_reduce_multi"]
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    lib.start_timing(names)
    for _ in range(args.parity_steps):
        step()
    torch.cuda.synchronize()
    timings = lib.stop_timing()
    conv_ms = sum(ms for evs in timings.values() for ms, _ in evs) / args.parity_steps
    model.enable_cuda_graphs(True)
    for _ in range(6):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.parity_steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.parity_steps
    tf = train_flops / (conv_ms / 1000.0) / 1e12
    out = {"precision": "fp32 activations, bf16x3 tcgen05 MMA (kind::f16, three terms, fp32 accumulate)",
           "value": batch / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms, "steps": args.parity_steps,
           "conv_ms_per_step": conv_ms, "conv_tflops_algorithmic": tf, "conv_frac_of_bf16_peak": tf / peak,
           "conv_tflops_executed": 3 * tf,
           "note": "the mode the 1e-4 parity tests (tests/test_gpu_fullsize.py, test_gpu_model.py) run in"}
    del model
    return out


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import numpy
    import torch
    import torch.distributed as dist
    from denet_b200 import lib
    from denet_b200.multi import ddp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback (use --impl reference)")
    # stdout carries exactly ONE line, the JSON result: everything else written to file descriptor 1 (NCCL prints its
    # version banner there when NCCL_DEBUG is set) is sent to stderr until that line is printed
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    rank, world = ddp.init_process_group()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    clib = lib.load()

    model, data_shape, batch, classes, solver = build_model(args.workload, args.batch, seed=1)
    model.to_device(dev, precision=args.precision)
    model.build_train_func(solver, [])
    if world > 1:
        model.enable_data_parallel()
    random.seed(1 + rank)
    x, metas = synthetic_batch(batch, data_shape, classes, seed=1 + rank)
    x_pinned = torch.from_numpy(x).pin_memory()
    x_dev = x_pinned.to(dev)
    hp = SOLVER_HP
    fwd_flops, train_flops, nconv = conv_flops(model)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    it = 0

    def step_device():
        nonlocal it
        model._train_step_device(x_dev, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"])
        it += 1

    def step_host():
        nonlocal it
        c = model.train_step(x_pinned, metas, 0, it, hp["lr"], hp["momentum"], hp["decay"])[0]
        it += 1
        return c

    # ---- conv kernel timing (roofline): every conv entry point bracketed by CUDA events on its stream.  With CUDA graphs
    # the events are EXTERNAL event-record nodes captured into the step's graphs and re-recorded by every replay, i.e. the
    # kernels are timed where they run in the timed regions below (an eager launch of a kernel shorter than the host's
    # ~20 us launch cost measures the host: the parity-class launches of the strided data gradients read 24 us eager,
    # 13 us as graph nodes); --eager times the eager launches.
    timed_names = ["denet_conv2d_fprop", "denet_conv2d_fprop_scatter", "denet_conv2d_dgrad_bnbwd", "denet_conv2d_wgrad",
                   "denet_conv2d_rowfold_fprop",
                   "denet_conv2d_rowfold_wgrad", "denet_wgrad_reduce_multi"]
    graphs = not args.eager
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    if not graphs:
        lib.start_timing(timed_names)
    le0 = clib.denet_launch_count()
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record()
    eager_steps = args.steps if not graphs else min(args.steps, 5)
    for _ in range(eager_steps):
        step_device()
    t1e.record()
    barrier()
    ms_eager = max_over_ranks(t0e.elapsed_time(t1e)) * args.steps / eager_steps
    launches = (clib.denet_launch_count() - le0) / eager_steps    # kernels per step (graph mode replays the same nodes)
    if not graphs:
        timings = lib.stop_timing()
    else:
        lib.start_timing(timed_names, in_graph=True)
        model.enable_cuda_graphs(True)
        for _ in range(6):                    # eager warm-up step, capture (events become graph nodes), first replays
            step_device()
        timings = {n: [] for n in timed_names}
        for _ in range(args.steps):
            step_device()
            for n, evs in lib.read_timing().items():
                timings[n].extend(evs)
        lib.stop_timing()
        barrier()

    graphs = not args.eager
    if graphs:
        model.enable_cuda_graphs(True)
    for _ in range(max(args.warmup, 3) + (3 if graphs else 0)):   # graph mode: eager warm-up, capture, then replays
                                                                  # (the first two replays of a fresh graph are slow)
        step_device()
    cost = step_host()
    if not numpy.isfinite(cost):
        raise SystemExit("bench.py: cost is not finite after warm-up (%r)" % cost)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()       # nvidia-smi stalls the GPU for ~10 ms while it attaches to the driver: run untimed steps
                              # until its first sample has arrived (it then samples every 100 ms through both regions)
    # ---- region A: inputs resident in HBM (value)
    for _ in range(8):        # untimed transition steps (input source host batch -> resident batch, sampler start-up);
        step_device()         # the same count on every rank (the steps contain collectives)
    extra = 0
    while world == 1 and sampler.proc is not None and len(sampler.rows) < 1 and extra < 50:
        step_device()
        extra += 1
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    marks = []
    for _ in range(args.steps):
        step_device()
        if args.detail:
            m = torch.cuda.Event(enable_timing=True)
            m.record()
            marks.append(m)
    e1.record()
    barrier()
    ms_a = max_over_ranks(e0.elapsed_time(e1))
    if args.detail and rank == 0:
        prev = e0
        per = []
        for m in marks:
            per.append(round(prev.elapsed_time(m), 2))
            prev = m
        print("region A per-step ms:", per, file=sys.stderr)

    # ---- region B: through the public API with HOST buffers: H2D of the batch and D2H of the costs every step
    barrier()
    from denet_b200 import layer as layer_mod
    for _ in range(3):        # untimed transition back to host batches
        cost = step_host()
    barrier()
    tb0 = dict(layer_mod.transfer_bytes)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    marks_b = []
    for _ in range(args.steps):
        cost = step_host()
        if args.detail:
            m = torch.cuda.Event(enable_timing=True)
            m.record()
            marks_b.append(m)
    e3.record()
    barrier()
    ms_b = max_over_ranks(e2.elapsed_time(e3))
    if args.detail and rank == 0:
        prev, per = e2, []
        for m in marks_b:
            per.append(round(prev.elapsed_time(m), 2))
            prev = m
        print("region B per-step ms:", per, file=sys.stderr)
    tb1 = dict(layer_mod.transfer_bytes)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        # leave together with rank 0; skip interpreter / NCCL teardown (captured graphs hold NCCL work: a one-sided
        # destroy_process_group hung the launcher for its whole timeout)
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)

    images = batch * world * args.steps
    value = images / (ms_a / 1000.0)
    e2e = images / (ms_b / 1000.0)

    # roofline of the dominant kernel family: the tcgen05 implicit-GEMM conv (fprop + dgrad share conv_fprop_kernel)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else \
        "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    per_layer = {}
    fam = {"fprop": [0.0, 0.0, 0], "dgrad": [0.0, 0.0, 0], "wgrad": [0.0, 0.0, 0]}     # ms, flops, launches
    for name, evs in timings.items():
        for ms, tag in evs:
            if name == "denet_wgrad_reduce_multi":
                fam["wgrad"][0] += ms          # the deferred split-K reduction of all filter gradients: time, no FLOPs
                continue
            if tag is None:
                continue
            kind, layer = tag
            k = "wgrad" if name.endswith("wgrad") else ("fprop" if kind == "fprop" else "dgrad")
            # a strided data gradient is stride_h * stride_w launches (one per parity class) that share the layer's FLOPs
            share = 1.0 / (layer.stride[0] * layer.stride[1]) if name.endswith("_scatter") else 1.0
            fam[k][0] += ms
            fam[k][1] += layer.fprop_flops() * share
            fam[k][2] += 1
            key = (k, layer.filter_shape, layer.input_shape, layer.stride)
            pl = per_layer.setdefault(key, [0.0, 0.0, 0])
            pl[0] += ms; pl[1] += layer.fprop_flops() * share; pl[2] += 1
    conv_ms = sum(v[0] for v in fam.values())
    conv_fl = sum(v[1] for v in fam.values())
    n_conv_launch = sum(v[2] for v in fam.values())
    achieved = conv_fl / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0
    # DRAM traffic per conv launch (dram__bytes_read.sum + dram__bytes_write.sum summed over the conv kernels of one
    # step / their launch count).  Hardware counters cannot be read inside a timed run: the figure comes from the `ncu`
    # launch list of this same command committed under profiles/ (newest round first) and is labelled with its source;
    # null when no such file exists.
    traffic, traffic_src = None, None
    for name in ("r2_conv_traffic_final.json", "r2_conv_traffic.json", "r1_conv_traffic_final.json"):
        try:
            traffic = float(json.load(open(os.path.join(ROOT, "profiles", name)))["conv_dram_bytes_per_launch"])
            traffic_src = "profiles/" + name
            break
        except (OSError, ValueError, KeyError):
            continue
    roofline = {"bound": "tensor", "kernel": "conv_fprop_halo_kernel / conv_fprop_kernel / conv_wgrad_kernel (+ split-K reduction) (tcgen05 implicit GEMM, bf16)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_unit": "DRAM bytes per conv launch; NOT measured in this run: ncu launch list of the same "
                                "command, %s" % traffic_src,
                "peak_source": peak_src,
                "flops_per_launch": conv_fl / max(n_conv_launch, 1), "ms_per_launch": conv_ms / max(n_conv_launch, 1),
                "launches_per_step": n_conv_launch / args.steps,
                "conv_ms_per_step": conv_ms / args.steps,
                "conv_share_of_step": (conv_ms / args.steps) / (ms_a / args.steps),
                "timed_in": ("external event-record nodes inside the step's CUDA graphs, read after each of %d replays "
                             "(eager launches: %.2f ms/step)" % (args.steps, ms_eager / args.steps)) if graphs else
                            "eager launches of the same steps (%.2f ms/step); value/e2e are eager too" % (
                                ms_eager / args.steps),
                "families": {k: {"tflops": (v[1] / (v[0] / 1000.0) / 1e12 if v[0] > 0 else None),
                                 "ms_per_step": v[0] / args.steps} for k, v in fam.items()},
                "algorithmic_gflop_per_image": train_flops / batch / 1e9}
    if args.detail:
        rows = sorted(per_layer.items(), key=lambda kv: -kv[1][0])
        for (k, fs, ishape, st), (ms, fl, n) in rows:
            print("%-6s filt %-20s in %-22s s%s  %8.3f ms/step  %7.1f TFLOP/s  x%d" % (
                k, fs, ishape, st[0], ms / args.steps, fl / (ms / 1000.0) / 1e12, n // args.steps), file=sys.stderr)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup + 1, "ms_per_step": ms_a / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": dict(workload_config(args.workload, batch, world, data_shape, classes, solver),
                           launch_mode="cuda-graphs" if graphs else "eager"),
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_b / args.steps,
                    "h2d_bytes_per_step": (tb1["h2d"] - tb0["h2d"]) // args.steps,
                    "d2h_bytes_per_step": (tb1["d2h"] - tb0["d2h"]) // args.steps},
            "gpu_launches": launches, "roofline": roofline, "last_cost": cost}
    if world == 1 and args.precision == "bf16" and not args.no_parity_mode:
        # SURVEY.md §7 / BASELINE.md §2 "report both": the same step in fp32-parity mode (fp32 activations, every conv
        # operand split into bf16 hi + lo, three MMA terms - the arithmetic the 1e-4 parity tests run in)
        del model
        torch.cuda.empty_cache()
        line["parity_mode"] = measure_parity_mode(args, dev, x_dev, metas, peak)
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_run(args.workload, args.cpu_sample_images, args.cpu_sample_steps, 1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    sys.stdout.flush()
    os.dup2(json_fd, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
