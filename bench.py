#!/usr/bin/env python
"""bench.py -- pileup windows/sec of the HELEN predict hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 256] [--features 10]
                    [--engine default|tensor|fp32] [--impl native|reference] [--sweep]

A *step* is one pass of the hot path over one batch of synthetic pileup windows
(uint8 [B, T=1000, F]) per GPU: all 19 chunks -> 2 x uint8[B, 1000] labels.

Printed JSON line (rank 0):
  value   windows/s, whole job (all ranks), inputs resident in HBM, CUDA-event timed per step
  e2e     same metric through the host-buffer entry (WindowPredictor.predict_host ->
          hb_predict_windows_host): pinned host images in, host labels out, copies in the timed region
  roofline      algorithmic FLOP/window x windows per launch / kernel time, vs measured bf16 peak
  cpu_baseline  the oracle's torch-CPU port of the reference predict loop on this box's host cores

--impl reference times that CPU port alone (the reference itself is a Python package that
cannot travel to the GPU box), same metric/config keys.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_COLUMNS = 1000
WINDOW, JUMP, HIDDEN = 100, 50, 128
L2_FLUSH_BYTES = 256 << 20


def flop_per_window(features, seq=T_COLUMNS):
    """Algorithmic FLOPs of one window: per chunk MACs = 76,800 F (enc ih)
    + 9,830,400 (enc hh) + 19,660,800 (dec ih) + 9,830,400 (dec hh) + 409,600 (heads)."""
    chunks = 0 if seq < WINDOW else (seq - WINDOW) // JUMP + 1
    mac = 2 * WINDOW * 3 * HIDDEN * features + 2 * (2 * WINDOW * 3 * HIDDEN * HIDDEN) \
        + 2 * WINDOW * 3 * HIDDEN * 2 * HIDDEN + WINDOW * 2 * HIDDEN * 16
    return 2 * mac * chunks


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"burst": float(p["bf16_tflops"]), "sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def kernel_traffic_bytes(engine, batch, features):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        table = json.load(f)
    return table.get(f"{engine}_B{batch}_F{features}")


def kernel_traffic_source():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get("source")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def random_parameters(features, seed):
    """Random-init TransducerGRU parameters (uniform +-1/sqrt(H), torch's GRU / Linear default range) under the reference's
    state_dict keys.  Own generator: the native arm imports nothing from oracle/."""
    import torch
    from collections import OrderedDict
    gen = torch.Generator().manual_seed(seed)
    bound = 1.0 / (HIDDEN ** 0.5)
    shapes = OrderedDict()
    for layer, k in (("gru_encoder", features), ("gru_decoder", 2 * HIDDEN)):
        for rev in ("", "_reverse"):
            shapes[f"{layer}.weight_ih_l0{rev}"] = (3 * HIDDEN, k)
            shapes[f"{layer}.weight_hh_l0{rev}"] = (3 * HIDDEN, HIDDEN)
            shapes[f"{layer}.bias_ih_l0{rev}"] = (3 * HIDDEN,)
            shapes[f"{layer}.bias_hh_l0{rev}"] = (3 * HIDDEN,)
    shapes["dense1_base.weight"], shapes["dense1_base.bias"] = (5, 2 * HIDDEN), (5,)
    shapes["dense2_rle.weight"], shapes["dense2_rle.bias"] = (11, 2 * HIDDEN), (11,)
    return OrderedDict((k, (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound) for k, shape in shapes.items())


def synthetic_images(batch, features, seed, seq=T_COLUMNS):
    import torch
    gen = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, seq, features), dtype=torch.uint8, generator=gen)


def cpu_port_windows_per_s(features, sample_windows, min_seconds, threads, seq=T_COLUMNS, reps_cap=64):
    """Times the oracle's torch-CPU port of the reference predict loop (the only place bench.py
    executes oracle/ code)."""
    import torch
    from oracle import TransducerPort, predict_port
    torch.set_num_threads(threads)
    model = TransducerPort(features).eval()
    model.load_state_dict(random_parameters(features, seed=0))
    images = synthetic_images(sample_windows, features, seed=1, seq=seq)
    predict_port(model, images[: max(1, sample_windows // 4)])          # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        predict_port(model, images)
        done += sample_windows
        dt = time.perf_counter() - t0
        if dt >= min_seconds or done >= reps_cap * sample_windows:
            break
    return done / dt, done, dt


def parity_sample(features, images_u8, base_labels, rle_labels, indices, margin=1e-5):
    """Checker leg (outside every timed region): the labels the GPU produced for `indices` of the timed batch against the
    oracle's torch-CPU port of predict.py:90-154 on the same images.  Flips are split by the oracle's own top-1/top-2
    margin (labels must agree wherever the margin is >= 1e-5; closer calls are counted, not gated)."""
    import numpy as np
    import torch
    from oracle import TransducerPort, predict_port
    from oracle.explicit import top2_margin
    model = TransducerPort(features).eval()
    model.load_state_dict(random_parameters(features, seed=0))
    ref = predict_port(model, images_u8[indices])
    out = {"windows": len(indices), "positions": int(2 * len(indices) * images_u8.shape[1]), "flips_above_margin": 0,
           "flips_sub_margin": 0, "margin": margin, "oracle": f"torch {torch.__version__} CPU fp32 port (oracle/torch_port.py)"}
    for got, ref_lab, ref_prob in ((base_labels, ref["base_label"], ref["base_prob"]), (rle_labels, ref["rle_label"], ref["rle_prob"])):
        diff = np.asarray(got)[indices] != ref_lab
        if diff.any():
            m = top2_margin(ref_prob)[diff]
            out["flips_sub_margin"] += int((m < margin).sum())
            out["flips_above_margin"] += int((m >= margin).sum())
    return out


def run_reference(args, rank, world):
    """--impl reference: the CPU arm alone.  Rank 0 only; other ranks exit quietly."""
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import TransducerPort, predict_port
    model = TransducerPort(args.features).eval()
    model.load_state_dict(random_parameters(args.features, seed=0))
    sample = args.reference_sample or args.batch              # same windows per step as the native arm
    images = synthetic_images(sample, args.features, seed=1000)
    for _ in range(args.warmup):
        predict_port(model, images[: max(1, sample // 8)])      # untimed: an eighth of a step each
    t0 = time.perf_counter()
    for _ in range(args.steps):
        predict_port(model, images)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "pileup windows/sec (B=256, T=1000)", "value": value, "unit": "windows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "windows/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x {sample} windows [T=1000, F={args.features}] per step "
                                   f"(warm-up steps: {max(1, sample // 8)} windows), "
                                   f"torch {torch.__version__} CPU fp32 nn.GRU port of the reference predict loop"},
        "probes": import_probes(),
        "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def import_probes():
    """Which of the reference's optional host-side dependencies exist on this box."""
    out = {}
    for name in ("h5py", "onnxruntime", "onnx"):
        try:
            mod = __import__(name)
            out[name] = getattr(mod, "__version__", "present")
        except Exception:
            out[name] = None
    return out


def workload_config(args, sample_windows=None):
    cfg = {"workload": f"1xB200 persistent bidir-GRU inference, B={args.batch}, T={T_COLUMNS}, F={args.features}, "
                       f"hidden=128, fused base+RLE heads (BASELINE configs[1])",
           "batch_per_gpu": args.batch, "seq_len": T_COLUMNS, "features": args.features,
           "chunk_width": WINDOW, "chunk_jump": JUMP, "chunks_per_window": 19,
           "sharding": f"windows x{args.gpus} ranks, no collective on the data path",
           "l2": "flushed between timed steps (256 MiB memset outside the per-step CUDA-event brackets)"}
    if sample_windows is not None:
        cfg["cpu_sample_windows_per_step"] = sample_windows
    return cfg


def run_native(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from helen_b200 import build as hb_build
    from helen_b200.predictor import WindowPredictor

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    if rank == 0:
        hb_build.build()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pred = WindowPredictor(random_parameters(args.features, seed=0), device=local_rank, engine=args.engine)
    engine = pred.engine
    host_images = synthetic_images(args.batch, args.features, seed=1000 + rank).pin_memory()   # e2e inputs start in pinned host memory
    images = host_images.to(dev)
    host_np = host_images.numpy()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def timed_steps(batch_images, steps, warmup, sampler=None):
        for _ in range(warmup):
            pred.predict(batch_images)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        barrier()
        if sampler:
            sampler.start()
        launches0 = pred.launch_count
        for i in range(steps):
            flush.zero_()
            starts[i].record()
            pred.predict(batch_images)
            ends[i].record()
        barrier()
        per_step = [s.elapsed_time(e) for s, e in zip(starts, ends)]
        return per_step, pred.launch_count - launches0

    sampler = ClockSampler(local_rank, period=0.02)
    pred.enable_kernel_timing(True)
    pred.kernel_time_ms(reset=True)
    per_step, launches = timed_steps(images, args.steps, args.warmup, sampler)
    clocks = sampler.finish()
    kernel_ms_total, kernel_calls = pred.kernel_time_ms(reset=True)
    pred.enable_kernel_timing(False)
    total_ms = max_over_ranks(sum(per_step))
    value = world * args.batch * args.steps / (total_ms * 1e-3)

    # dominant kernel (GRU recurrence) timed on its own: separate short pass, CUDA events around every launch
    dom_ms, dom_launches = None, 0
    if engine == "tensor":
        pred.enable_kernel_timing(2)
        pred.dominant_kernel_time_ms(reset=True)
        for _ in range(3):
            flush.zero_()
            pred.predict(images)
        torch.cuda.synchronize()
        dom_ms, dom_launches = pred.dominant_kernel_time_ms(reset=True)
        pred.enable_kernel_timing(False)
    pred.predict(images)
    torch.cuda.synchronize()
    plan = pred.last_launch_plan()

    # end to end through the host-buffer entry: pinned host images in (numpy view), numpy labels out
    for _ in range(max(1, args.warmup // 2)):
        pred.predict_host(host_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pred.predict_host(host_np)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * args.batch * args.steps / e2e_s

    # sustained: back-to-back steps for >= --sustained-seconds (the K timed steps above last tens of milliseconds),
    # L2 flushed before every step as above; one CUDA-event bracket around the whole run, clocks sampled during it
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds / max(sum(per_step) / len(per_step) * 1e-3, 1e-6)))
        sus_sampler = ClockSampler(local_rank, period=0.05)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sus_sampler.start()
        ev0.record()
        for _ in range(n_sus):
            flush.zero_()
            pred.predict(images)
        ev1.record()
        barrier()
        sus_ms = max_over_ranks(ev0.elapsed_time(ev1))
        sustained = {"value": world * args.batch * n_sus / (sus_ms * 1e-3), "unit": "windows/s", "steps": n_sus,
                     "seconds": sus_ms * 1e-3, "includes": "the 256 MiB L2 flush before every step",
                     "clocks": sus_sampler.finish()}

    # labels of the timed batch for the parity block (checked against the oracle below, outside every timed region)
    chk_base, chk_rle = pred.predict(images)
    torch.cuda.synchronize()
    chk_base, chk_rle = chk_base.cpu().numpy(), chk_rle.cpu().numpy()

    sweep = None
    if args.sweep and world == 1:
        sweep = []
        for b in (64, 128, 256, 512, 1024, 2048):
            imgs = synthetic_images(b, args.features, seed=7).to(dev)
            ps, _ = timed_steps(imgs, max(3, args.steps // 4), 2)
            sweep.append({"batch": b, "windows_per_s": b * len(ps) / (sum(ps) * 1e-3), "ms_per_step": sum(ps) / len(ps)})

    if rank != 0:
        return
    peaks = measured_peaks()
    fpw = flop_per_window(args.features)
    # the dominant kernel covers a whole batch per launch sequence; its device time comes from the
    # library's own event bracket around that sequence on the launching stream
    kernel_ms = kernel_ms_total / max(kernel_calls, 1) if kernel_calls else statistics.mean(per_step)
    achieved_tf = args.batch * fpw / (kernel_ms * 1e-3) / 1e12
    path_roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["burst"], "unit": "TFLOP/s",
                     "frac": achieved_tf / peaks["burst"], "frac_of_sustained": achieved_tf / peaks["sustained"],
                     "flop_per_window": fpw, "ms_per_batch": kernel_ms,
                     "what": "all kernels of one batch (library event bracket on the launching stream)"}
    if dom_launches and plan["chunkloop"]:
        # one launch = the whole chunk loop of the batch: per chunk enc hh + dec ih + dec hh + heads (SURVEY 8d terms;
        # the encoder input projection runs once per batch in its own kernel and is not counted here)
        chunks = (T_COLUMNS - WINDOW) // JUMP + 1
        mac_chunk = 2 * (2 * WINDOW * 3 * HIDDEN * HIDDEN) + 2 * WINDOW * 3 * HIDDEN * 2 * HIDDEN + WINDOW * 2 * HIDDEN * 16
        dom_flop = args.batch * chunks * 2 * mac_chunk
        dom_ms_launch = dom_ms / dom_launches
        dom_tf = dom_flop / (dom_ms_launch * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "tc_chunkloop_kernel" if plan["windows_per_cta"] == 8 else "tc_chunkloop2_kernel", "achieved": dom_tf, "peak": peaks["burst"],
                    "unit": "TFLOP/s", "frac": dom_tf / peaks["burst"], "frac_of_sustained": dom_tf / peaks["sustained"],
                    "peak_source": ("measured" if peaks["source"] == "measured" else "fallback") + " bf16 burst",
                    "flop_per_launch": dom_flop, "kernel_ms_per_launch": dom_ms_launch, "launches_timed": dom_launches,
                    "us_per_dependent_step": 1e3 * dom_ms_launch / (chunks * 2 * WINDOW),
                    "traffic": kernel_traffic_bytes(engine, args.batch, args.features),
                    "traffic_source": kernel_traffic_source(),
                    "note": "latency-bound at this batch: 3,800 dependent GRU steps per launch (DESIGN.md section 5)"}
        # executed (not algorithmic) MMA work: the fp16 split runs the recurrence as 4 partial products when the
        # [h_hi | h_lo] operand is stacked (3 otherwise), the decoder projection and the heads as 3
        rec_terms = 4 if plan["stacked_operand"] else 3
        mac_exec = rec_terms * 2 * (2 * WINDOW * 3 * HIDDEN * HIDDEN) + 3 * (2 * WINDOW * 3 * HIDDEN * 2 * HIDDEN) + 3 * (WINDOW * 2 * HIDDEN * 16)
        roofline["executed_over_algorithmic"] = mac_exec / mac_chunk
        roofline["frac_executed"] = roofline["frac"] * mac_exec / mac_chunk
    elif dom_launches:
        # one recurrence launch = batch windows x 100 dependent steps x 2 directions of one layer
        rec_flop = args.batch * WINDOW * 2 * (2 * 3 * HIDDEN * HIDDEN)
        rec_ms = dom_ms / dom_launches
        rec_tf = rec_flop / (rec_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "tc_recurrence_kernel", "achieved": rec_tf, "peak": peaks["burst"],
                    "unit": "TFLOP/s", "frac": rec_tf / peaks["burst"], "frac_of_sustained": rec_tf / peaks["sustained"],
                    "peak_source": ("measured" if peaks["source"] == "measured" else "fallback") + " bf16 burst",
                    "flop_per_launch": rec_flop, "kernel_ms_per_launch": rec_ms, "launches_timed": dom_launches,
                    "us_per_dependent_step": 1e3 * rec_ms / WINDOW,
                    "traffic": kernel_traffic_bytes(engine, args.batch, args.features),
                    "note": "latency-bound at this batch: 100 dependent GRU steps per launch (DESIGN.md section 5)"}
    else:
        roofline = dict(path_roofline, traffic=kernel_traffic_bytes(engine, args.batch, args.features))
    line = {
        "metric": "pileup windows/sec (B=256, T=1000)", "value": value, "unit": "windows/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if engine == "fp32" else "f16x3-split operands, f32 accumulate/state",
        "data": "synthetic", "engine": engine, "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": "windows/s", "h2d_bytes_per_step": int(host_np.nbytes),
                "d2h_bytes_per_step": int(2 * args.batch * T_COLUMNS)},
        "gpu_launches": int(launches),
        "launch_plan": plan,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_path": path_roofline,
    }
    if sustained:
        line["sustained"] = sustained
    if sweep:
        line["batch_sweep"] = sweep
    line["probes"] = import_probes()
    if not args.no_parity:
        idx = sorted(set(int(i) for i in __import__("numpy").linspace(0, args.batch - 1, min(args.parity_windows, args.batch)).astype(int)))
        line["parity"] = parity_sample(args.features, host_images, chk_base, chk_rle, idx)
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, done, dt = cpu_port_windows_per_s(args.features, args.cpu_sample, args.cpu_seconds, threads)
        line["cpu_baseline"] = {"value": v, "unit": "windows/s", "cores": threads, "kind": "port",
                                "sample": f"{done} windows [T=1000, F={args.features}] in batches of {args.cpu_sample}, "
                                          f"{dt:.1f} s, torch {torch.__version__} CPU fp32 nn.GRU port of predict.py:90-154"}
    print(json.dumps(line), flush=True)
    pred.close()
    if world > 1:
        dist.destroy_process_group()


def run_polish(args, rank, world, local_rank):
    """--polish N: BASELINE configs[2], window-sharded polish of N synthetic windows (the 3 Gb contig set is ~3 M): every rank
    generates its contiguous shard on the device in batches of --batch windows (seed 1000 + rank) and predicts it.  The
    "stitch" of this mode is a HOST-side ordered gather with no collective: the stitched result lives in one shared host
    array (POSIX shared memory, page-locked in every rank), and each rank copies the uint8 labels of every batch straight
    from its GPU into its own window range of that array, asynchronously on a copy stream while the next batch runs
    (two device label buffers in rotation).  Secondary line; the parity block checks a few windows per rank against the oracle."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from helen_b200 import build as hb_build
    from helen_b200.predictor import WindowPredictor
    from helen_b200.sharding import shard_bounds
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --polish: no CUDA device; there is no CPU fallback")
    if rank == 0:
        hb_build.build()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pred = WindowPredictor(random_parameters(args.features, seed=0), device=local_rank, engine=args.engine)
    start, end = shard_bounds(args.polish, world, rank)
    n_local = end - start
    # the stitched result: [N windows, 2 heads, T] uint8 in shared host memory, created by rank 0
    shm_path = f"/dev/shm/helen_b200_polish_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}"
    total_bytes = 2 * args.polish * T_COLUMNS
    if rank == 0:
        with open(shm_path, "wb") as f:
            f.truncate(total_bytes)
    if world > 1:
        dist.barrier()
    # layout [window, head, T]: a rank's window range is one contiguous piece of the array, the only piece it page-locks
    host = torch.from_file(shm_path, shared=True, size=total_bytes, dtype=torch.uint8).view(args.polish, 2, T_COLUMNS)
    cudart = torch.cuda.cudart()
    mine = host[start:end]
    registered = False
    if n_local > 0:
        try:
            registered = int(cudart.cudaHostRegister(mine.data_ptr(), mine.numel(), 0)) == 0     # page-locked: async DMA straight into it
        except Exception:
            registered = False
        if not registered:                                 # (locked-memory limit): clear the error, the copies below then stage through the driver
            try:
                import ctypes
                ctypes.CDLL("libcudart.so.12").cudaGetLastError()
            except OSError:
                pass
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    warm = torch.randint(0, 256, (min(args.batch, max(n_local, 1)), T_COLUMNS, args.features), dtype=torch.uint8, device=dev, generator=gen)
    for _ in range(3):
        pred.predict(warm)
    copy_stream = torch.cuda.Stream(device=dev)
    done = [torch.cuda.Event(), torch.cuda.Event()]            # label buffer b may be overwritten once its copies have run
    keep = min(args.parity_windows, max(n_local, 1))             # images of the first few windows, for the oracle check
    kept_images = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    main = torch.cuda.current_stream(dev)
    bufs = [None, None]
    for it, off in enumerate(range(0, n_local, args.batch)):
        n = min(args.batch, n_local - off)
        images = torch.randint(0, 256, (n, T_COLUMNS, args.features), dtype=torch.uint8, device=dev, generator=gen)
        if it == 0:
            kept_images = images[:keep].clone()
        if it >= 2:
            main.wait_event(done[it & 1])
        b, r = pred.predict(images)
        bufs[it & 1] = (b, r)                                    # keep the buffers alive until their copies have run
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            host[start + off:start + off + n, 0].copy_(b, non_blocking=True)
            host[start + off:start + off + n, 1].copy_(r, non_blocking=True)
            done[it & 1].record(copy_stream)
    main.synchronize()
    t_compute = time.perf_counter() - t0
    copy_stream.synchronize()
    if world > 1:
        dist.barrier()                                           # every rank's range of the shared array is complete
    t_total = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_compute, t_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_compute, t_total = t.tolist()
    # checker leg (untimed): this rank's first windows against the oracle, read back from the stitched array
    flips = torch.zeros(3, dtype=torch.int64, device=dev)
    if not args.no_parity and n_local > 0:
        chk = parity_sample(args.features, kept_images.cpu(), host[start:start + keep, 0].numpy(), host[start:start + keep, 1].numpy(),
                            list(range(keep)))
        flips = torch.tensor([chk["windows"], chk["flips_above_margin"], chk["flips_sub_margin"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(flips)
    if rank == 0:
        checksum = int(host[::997].to(torch.int64).sum())
        print(json.dumps({
            "metric": "window-sharded polish, end-to-end windows/sec (generate on device, predict, labels streamed into one host array)",
            "value": args.polish / t_total, "unit": "windows/s", "n_gpus": world, "windows": args.polish, "batch_per_launch": args.batch,
            "seconds": t_total, "predict_seconds": t_compute, "predict_windows_per_s": args.polish / t_compute,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16x3-split operands, f32 accumulate/state",
            "data": "synthetic", "label_checksum": checksum, "host_array_page_locked": registered,
            "parity": {"windows": int(flips[0]), "flips_above_margin": int(flips[1]), "flips_sub_margin": int(flips[2]),
                       "what": f"first {keep} windows of every rank's shard against the oracle port, read from the stitched host array"},
            "config": {"workload": f"{world}xB200 window-sharded polish over {args.polish} synthetic windows [T=1000, F={args.features}] "
                                   f"in batches of {args.batch}, host-side ordered gather through shared host memory, no collective "
                                   f"(BASELINE configs[2])"},
        }), flush=True)
    if registered:
        cudart.cudaHostUnregister(mine.data_ptr())
    del mine, host
    if world > 1:
        dist.barrier()
    if rank == 0:
        os.unlink(shm_path)
    pred.close()
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """--train: BASELINE configs[3], one chunk step = forward + CE(base) + weighted CE(rle) + backward on [B=128, 100, F]
    (train.py:189-201 through ChunkTrainer.step -> hb_train_step_chunk); secondary line, not the headline metric."""
    import torch
    from helen_b200 import build as hb_build
    from helen_b200.models.TransducerModel import TransducerGRU
    from helen_b200.models.train_step import ChunkTrainer
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --train: no CUDA device; there is no CPU fallback")
    hb_build.build()
    batch, features = args.train_batch, args.features
    sd = random_parameters(features, seed=0)
    model = TransducerGRU(1, features, 1, HIDDEN, 5, 11)
    model.load_state_dict(sd)
    model = model.cuda()
    trainer = ChunkTrainer(model)
    gen = torch.Generator().manual_seed(2)
    x = torch.randint(0, 256, (batch, WINDOW, features), generator=gen).float().cuda()
    lb = torch.randint(0, 5, (batch, WINDOW), generator=gen).cuda()
    lr = torch.randint(0, 11, (batch, WINDOW), generator=gen).cuda()
    hidden = None
    for _ in range(args.warmup):
        _, _, _, hidden = trainer.step(x, hidden, lb, lr)
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        loss, _, _, hidden = trainer.step(x, hidden, lb, lr)
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / args.steps
    fwd_mac = 2 * WINDOW * 3 * HIDDEN * features + 2 * (2 * WINDOW * 3 * HIDDEN * HIDDEN) + 2 * WINDOW * 3 * HIDDEN * 2 * HIDDEN + WINDOW * 2 * HIDDEN * 16
    flop = 3 * 2 * fwd_mac * batch                               # backward ~ 2x forward
    props = torch.cuda.get_device_properties(0)
    fp32_peak = props.multi_processor_count * 128 * 2 * 1.965e9 / 1e12
    # CPU arm (cpu_baseline leg: the one place this mode executes oracle/): the port's autograd with the reference's criteria
    from oracle import TransducerPort
    from helen_b200.options import TrainOptions
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    port = TransducerPort(features)
    port.load_state_dict(sd)
    crit_b = torch.nn.CrossEntropyLoss()
    crit_r = torch.nn.CrossEntropyLoss(weight=torch.tensor(TrainOptions.CLASS_WEIGHTS))
    xc, lbc, lrc = x.cpu(), lb.cpu(), lr.cpu()
    hc = torch.zeros(batch, 2, HIDDEN)
    cpu_steps, t0 = 0, time.perf_counter()
    while cpu_steps < 2 or time.perf_counter() - t0 < 5.0:
        port.zero_grad()
        ob, orl, hc = port(xc, hc)
        (crit_b(ob.reshape(-1, 5), lbc.reshape(-1)) + crit_r(orl.reshape(-1, 11), lrc.reshape(-1))).backward()
        hc = hc.detach()
        cpu_steps += 1
    cpu_rate = cpu_steps / (time.perf_counter() - t0)
    print(json.dumps({
        "metric": "train chunk-steps/sec (B=128, W=100; forward + CE + weighted CE + backward)", "value": 1e3 / ms, "unit": "chunk-steps/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "last_loss": loss,
        "config": {"workload": f"helen_train forward+backward GRU on 1xB200, B={batch}, W={WINDOW}, F={features} (BASELINE configs[3])"},
        "roofline": {"bound": "fp32 FMA", "achieved": flop / (ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": flop / (ms * 1e-3) / 1e12 / fp32_peak, "peak_source": "computed: SMs x 128 lanes x 2 x 1.965 GHz (not measured)",
                     "flop_per_step": flop, "traffic": None},
        "cpu_baseline": {"value": cpu_rate, "unit": "chunk-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{cpu_steps} steps, torch {torch.__version__} CPU autograd of the nn.GRU port with the reference's criteria"},
    }), flush=True)
    trainer.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="windows per GPU per step")
    ap.add_argument("--features", type=int, default=10)
    ap.add_argument("--engine", default="default", choices=["default", "tensor", "fp32"])
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--sweep", action="store_true", help="also report the batch sweep 64..2048 (BASELINE configs[4])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of a sample of the timed batch")
    ap.add_argument("--parity-windows", type=int, default=16, help="windows of the timed batch checked against the oracle")
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the extra back-to-back run (0 = skip)")
    ap.add_argument("--cpu-sample", type=int, default=64, help="windows per CPU-baseline batch")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--reference-sample", type=int, default=0, help="windows per step of --impl reference (0 = --batch, the native arm's step)")
    ap.add_argument("--polish", type=int, default=0, metavar="N",
                    help="secondary line: BASELINE configs[2], window-sharded polish of N synthetic windows (use --batch 2048)")
    ap.add_argument("--train", action="store_true", help="secondary line: BASELINE configs[3], one training chunk step at B=128")
    ap.add_argument("--train-batch", type=int, default=128)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world != args.gpus and world == 1 and args.gpus > 1 and args.impl == "native" and not args.train:
        # launched without torchrun: re-exec under torch.distributed.run
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"),
               os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.polish:
        run_polish(args, rank, world, local_rank)
    elif args.train:
        if rank == 0:
            run_train(args)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
