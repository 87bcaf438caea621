#!/usr/bin/env python
"""bench.py — activation rows/s of the TopK-SAE train step (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 50 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # CPU arm: the oracle port on the host cores

Workload (BASELINE.json configs[1]): whisper-tiny TopKSAE 384 -> 3072, k = 32, hyper-parameters of
configs/tiny_default.yaml (lr 1e-4, warm-up 1000, clip 1.0, AMP on => bf16 path), synthetic
row-standardised Gaussian activations.  One "step" = one `SAETrainer.train_step`
(pack, tcgen05 GEMM + fused TopK, sparse decode + MSE, sparse backward, clip + AdamW, decoder
renorm, counters, stats readback).  The YAML batch (128) is launch-latency bound by construction,
so the headline batch is `--batch` (default 75776 = 4 waves of 148 x 128-row tiles; config-legal:
TrainingConfig.batch_size >= 1);
the YAML-batch number is reported next to it as `yaml_batch`.

N > 1: one process per GPU, each training an independent layer's SAE (4 encoder + 4 decoder
layers of whisper-tiny; no data-path collective), value = sum of rows/s, weak scaling.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# Workloads (BASELINE.json configs): "tiny" = configs[1] (headline; N GPUs train N independent layer
# SAEs, no collective); "small-dp" / "large-dp" = configs[2] / [3]: one SAE, global batch sharded over
# the ranks, gradients all-reduced with NCCL (strong scaling).
WORKLOADS = {
    "tiny": dict(label="whisper-tiny", d=384, expansion=8, k=32, dp=False, cfg="configs[1]"),
    "small-dp": dict(label="whisper-small", d=768, expansion=8, k=32, dp=True, cfg="configs[2]"),
    "large-dp": dict(label="whisper-large-v3", d=1280, expansion=32, k=32, dp=True, cfg="configs[3]"),
}
D_MODEL, EXPANSION, TOPK = 384, 8, 32
HIDDEN = D_MODEL * EXPANSION


def set_workload(name: str) -> dict:
    global D_MODEL, EXPANSION, TOPK, HIDDEN
    w = WORKLOADS[name]
    D_MODEL, EXPANSION, TOPK = w["d"], w["expansion"], w["k"]
    HIDDEN = D_MODEL * EXPANSION
    return w
METRIC = "sae_train_step_activation_rows_per_sec"
UNIT = "rows/s"


def load_peaks() -> tuple[dict, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: an NVML polling thread (5 ms
    cadence, so even a 30 ms region gets several samples); falls back to `nvidia-smi -lms`."""

    _REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.samples: list[tuple[float, float, int]] = []
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self.proc = None
        self.lines: list[str] = []

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.gpu_index < len(ids) and ids[self.gpu_index].isdigit():
                return int(ids[self.gpu_index])
        return self.gpu_index

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self._nvml = pynvml
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self._stop.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        try:
                            rs = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        except Exception:
                            rs = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.samples.append((sm, mx, rs))
                    except Exception:
                        pass
                    self._stop.wait(0.005)

            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None
            try:
                fields = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                          "clocks_event_reasons.hw_thermal_slowdown,"
                          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.proc = subprocess.Popen(
                    ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={fields}",
                     "--format=csv,noheader,nounits", "-lms", "100"],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self._thread = threading.Thread(target=self._pump, daemon=True)
                self._thread.start()
            except OSError:
                self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self) -> dict:
        if self.samples:
            sm = [x[0] for x in self.samples]
            reasons = set()
            for _, _, rs in self.samples:
                for name, bit in self._REASONS:
                    if rs & bit:
                        reasons.add(name)
            return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm),
                    "sm_max_mhz": self.samples[0][1], "reasons": sorted(reasons),
                    "samples": len(sm), "source": "nvml"}
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def make_trainer(batch: int, device: str, layer_seed: int, use_amp: bool = True, cuda_graph=None,
                 data_parallel: bool = False):
    from whisper_sae_b200.config import ExperimentConfig
    from whisper_sae_b200.sae import SAETrainer, create_sae

    cfg = ExperimentConfig.from_yaml(ROOT / "configs" / "tiny_default.yaml")
    cfg.training.batch_size = batch
    cfg.training.use_amp = use_amp
    cfg.sae.expansion_factor = EXPANSION
    cfg.sae.k = TOPK
    torch.manual_seed(cfg.training.seed + layer_seed)
    sae = create_sae(cfg.sae, D_MODEL)
    run_dir = Path(tempfile.mkdtemp(prefix="wsae_bench_"))
    tr = SAETrainer(sae, cfg.training, device=device, run_dir=run_dir, cuda_graph=cuda_graph,
                    data_parallel=data_parallel)
    tr.setup_scheduler(100_000)
    return tr, cfg


def synth(n_rows: int, d: int, seed: int, device=None, pin: bool = False) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_rows, d, generator=g)
    x = (x - x.mean(1, keepdim=True)) / x.std(1, unbiased=False, keepdim=True)
    if device is not None:
        return x.to(device)
    return x.pin_memory() if pin else x


def time_steps(tr, batches, steps: int, warmup: int, dist_on: bool) -> tuple[float, int]:
    """(seconds, own kernel launches) for exactly `steps` train steps: CUDA events on the launching
    stream, barrier + synchronize on both sides, max over ranks."""
    from whisper_sae_b200 import ops

    n = len(batches)
    for i in range(warmup):
        tr.train_step(batches[i % n])
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ops.GPU_LAUNCHES
    start.record()
    for i in range(steps):
        tr.train_step(batches[(warmup + i) % n])
    end.record()
    torch.cuda.synchronize()
    launches = ops.GPU_LAUNCHES - launches0
    if dist_on:
        torch.distributed.barrier()
    sec = start.elapsed_time(end) / 1e3
    if dist_on:
        t = torch.tensor([sec], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        sec = t.item()
    return sec, launches


def time_epoch(tr, host_batches, steps: int, warmup: int, dist_on: bool) -> float:
    """End-to-end seconds for `steps` steps through the public loop `SAETrainer.train_epoch` fed with
    pinned HOST batches: every step's H2D copy (prefetched one batch ahead on a copy stream) and
    its metrics readback (32 bytes the counters kernel posts to the pinned host mailbox: sse, l0
    count, dead count, sequence word) are inside the timed region."""
    n = len(host_batches)
    tr.train_epoch([[host_batches[i % n]] for i in range(warmup)])
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    feed = [[host_batches[(warmup + i) % n]] for i in range(steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    tr.train_epoch(feed)
    end.record()
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
    sec = start.elapsed_time(end) / 1e3
    if dist_on:
        t = torch.tensor([sec], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        sec = t.item()
    return sec


def kernel_profile(tr, batches, steps: int) -> dict:
    from whisper_sae_b200 import ops

    ops.PROFILE = ops.KernelProfile()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        tr.train_step(batches[i % len(batches)])
    t1.record()
    summ = ops.PROFILE.summary()
    ops.PROFILE = None
    total_ms = t0.elapsed_time(t1)
    for v in summ.values():
        v["share_of_step"] = v["total_ms"] / total_ms
        v["per_step"] = v["launches"] / steps
    summ["_step_ms"] = total_ms / steps
    return summ


def roofline(prof: dict, batch: int, d: int, F: int, k: int, peaks: dict, bf16_dec: bool) -> dict:
    """Roofline of the dominant kernel (largest share of the step), DESIGN.md §kernels:
    encode_topk: tensor bound, 2*B*d*F algorithmic FLOPs per launch;
    decode: HBM bound, B*(k*d*w + 2*d*4 + k*8) bytes; backward: B*(k*d*w + 2*2*k*d*4 + d*4 + k*12)."""
    kern = {n: v for n, v in prof.items() if not n.startswith("_")}
    top = max(kern, key=lambda n: kern[n]["total_ms"])
    per_launch_s = kern[top]["avg_ms"] / 1e3
    w = 2 if bf16_dec else 4
    if top in ("wsae_encode_topk", "wsae_wgrad_gemm"):
        work = 2.0 * batch * d * F
        peak = peaks["bf16_tflops_sustained"]
        return {"kernel": top, "bound": "tensor", "achieved": work / per_launch_s / 1e12, "peak": peak,
                "unit": "TFLOP/s", "frac": work / per_launch_s / 1e12 / peak, "traffic": None}
    if top == "wsae_decode_mse":
        nbytes = batch * (k * d * w + 2 * d * 4 + k * 8)
    elif top == "wsae_decode_backward":   # gathered rows once + x read + bf16 residual write + idx/val/dv
        nbytes = batch * (k * d * w + d * 4 + d * 2 + k * 12)
    elif top == "wsae_backward_sparse":
        nbytes = batch * (k * d * w + 2 * 2 * k * d * 4 + 2 * d * 4 + k * 12)
    elif top == "wsae_fused_adamw":
        nbytes = 28 * (kern[top]["launches"] and (2 * d * F + F + 2 * d)) / 5  # avg per launch (5 tensors)
    else:
        nbytes = 0
    peak = peaks["hbm_gbs"]
    ach = nbytes / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
    return {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "traffic": None}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` of the default
# workload (tiny, B = 75776, bf16): profiles/r1_v12_top3_ncu_full.txt.  Other shapes: not captured.
NCU_DRAM_TRAFFIC = {("tiny", 75776): {"wsae_encode_topk": 74.24e6, "wsae_decode_backward": 184.51e6,
                                      "wsae_wgrad_gemm": 90.85e6}}


def cpu_oracle_rate(batch: int, seconds_budget: float, threads: int) -> dict:
    """rows/s of the CPU oracle port (reference algorithm, torch CPU fp32, all host threads)."""
    from oracle import topk_sae_oracle as O

    torch.set_num_threads(threads)
    torch.manual_seed(42)
    state = O.init_state(D_MODEL, HIDDEN)
    opt = O.AdamWState()
    lrs = O.lr_sequence(1e-4, 100_000, 1000)
    x = O.synthetic_activations(batch * 2, D_MODEL, seed=1234)
    O.train_step(state, opt, x[:batch], TOPK, lrs[0])          # warm-up (thread pools, page faults)
    n, t0 = 0, time.perf_counter()
    while True:
        O.train_step(state, opt, x[(n % 2) * batch:(n % 2 + 1) * batch], TOPK, lrs[n + 1])
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds_budget or n >= 200:
            break
    return {"value": batch * n / el, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} train steps of B={batch} rows ({el:.1f} s), oracle/topk_sae_oracle.py, "
                      f"torch {torch.__version__} CPU fp32"}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps, warm = max(1, args.steps), max(0, args.warmup)
    from oracle import topk_sae_oracle as O

    torch.set_num_threads(threads)
    torch.manual_seed(42)
    state = O.init_state(D_MODEL, HIDDEN)
    opt = O.AdamWState()
    lrs = O.lr_sequence(1e-4, 100_000, 1000)
    batch = args.batch
    x = O.synthetic_activations(batch * 2, D_MODEL, seed=1234)
    # bounded: at most ~120 s of CPU work in total
    probe0 = time.perf_counter()
    O.train_step(state, opt, x[:batch], TOPK, lrs[0])
    per = time.perf_counter() - probe0
    max_steps = max(1, int(120.0 / max(per, 1e-3)))
    warm = min(warm, max(0, max_steps // 4))
    steps = min(steps, max(1, max_steps - warm))
    for i in range(warm):
        O.train_step(state, opt, x[(i % 2) * batch:(i % 2 + 1) * batch], TOPK, lrs[i + 1])
    t0 = time.perf_counter()
    for i in range(steps):
        O.train_step(state, opt, x[(i % 2) * batch:(i % 2 + 1) * batch], TOPK, lrs[warm + i + 1])
    el = time.perf_counter() - t0
    value = batch * steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * el / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same config block as the B200 arm at this N; the host trains the N per-GPU units one after
        # another (reference scripts/train.py:338-342 loops the layers), so its rows/s does not depend on N
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} timed train steps of B={batch} rows on the host CPU "
                                   f"(oracle port of sae/model.py + sae/training.py, torch CPU fp32; "
                                   f"units of an N-GPU job run sequentially at this rate)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n_gpus: int) -> dict:
    w = WORKLOADS[args.workload]
    if w["dp"] and n_gpus > 1:
        par = f"batch sharded over {n_gpus} GPUs, NCCL all-reduce of the gradient bucket"
    elif n_gpus > 1:
        par = "1 SAE (layer) per GPU, no data-path collective"
    else:
        par = "single GPU"
    return {
        "workload": f"{w['label']} TopKSAE {D_MODEL}->{HIDDEN} k={TOPK} train step "
                    f"(configs/tiny_default.yaml hyper-parameters; BASELINE.json {w['cfg']})",
        "batch_rows_per_gpu": args.batch,
        "global_batch_rows": args.batch * n_gpus,
        "parallelism": par,
        "precision_mode": args.precision,
        "l2_policy": f"{args.resident_batches} distinct resident batches cycled "
                     f"({args.resident_batches * args.batch * D_MODEL * 4 / 2**20:.0f} MiB > 126 MiB L2)",
    }


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="tiny")
    ap.add_argument("--batch", type=int, default=None,
                    help="rows per GPU per step (default 75776 = 148 SMs x 128 rows x 4; dp workloads: "
                         "that global batch / N)")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--resident-batches", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--value-only", action="store_true",
                    help="only the HBM-resident timed leg (used under ncu; prints a reduced line)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    wl = set_workload(args.workload)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.batch is None:
        # K1 walks the batch in 128-row tiles on 148 persistent CTAs: 148 * 128 * 4 = 75776 rows is
        # exactly four waves (65536 would be 3.46 -> 4 waves with a 14 % idle tail)
        args.batch = 75776 // world_env if wl["dp"] else 75776
        if args.workload == "large-dp":
            args.batch = 37888 // world_env

    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    if dist_on:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = f"cuda:{local_rank}"

    from whisper_sae_b200 import ops

    peaks, peak_src = load_peaks()
    dp = wl["dp"] and dist_on
    tr, cfg = make_trainer(args.batch, dev, layer_seed=0 if dp else rank,
                           use_amp=(args.precision == "bf16"), data_parallel=dp)
    nb = max(2, args.resident_batches)
    rows = synth(nb * args.batch, D_MODEL, seed=1234 + rank)
    dev_batches = [rows[i * args.batch:(i + 1) * args.batch].to(dev) for i in range(nb)]
    host_batches = [rows[i * args.batch:(i + 1) * args.batch].pin_memory() for i in range(nb)]

    # ---- value: inputs resident in HBM ----
    with ClockSampler(local_rank) as clocks:
        sec, launches = time_steps(tr, dev_batches, args.steps, args.warmup, dist_on)
    value = args.batch * args.steps * world / sec

    if args.value_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                              "steps": args.steps, "warmup": args.warmup,
                              "ms_per_step": 1e3 * sec / args.steps, "gpu_launches": launches,
                              "note": "value-only leg (not a bench line)"}))
        if dist_on:
            torch.distributed.destroy_process_group()
        return

    # ---- e2e: pinned host batches through SAETrainer.train_step (H2D + stats D2H inside) ----
    sec_e2e = time_epoch(tr, host_batches, args.steps, args.warmup, dist_on)
    e2e = args.batch * args.steps * world / sec_e2e

    line = None
    if rank == 0:
        # per-kernel CUDA-event timing needs eager launches: same kernels, graph replay switched off
        tr_eager, _ = make_trainer(args.batch, dev, layer_seed=rank, use_amp=(args.precision == "bf16"),
                                   cuda_graph="eager")       # single-rank kernels (no collective)
        for i in range(3):
            tr_eager.train_step(dev_batches[i % nb])
        prof = kernel_profile(tr_eager, dev_batches, min(args.steps, 20))
        del tr_eager
        roof = roofline(prof, args.batch, D_MODEL, HIDDEN, TOPK, peaks, args.precision == "bf16")
        roof["peak_source"] = f"{peak_src} (MEASURED_PEAKS.json)" if peak_src == "measured" else "fallback"
        roof["traffic"] = NCU_DRAM_TRAFFIC.get((args.workload, args.batch), {}).get(roof["kernel"])
        roof["traffic_source"] = "profiles/r1_v12_top3_ncu_full.txt (dram bytes per launch)" \
            if roof["traffic"] is not None else None
        shares = {n: round(v["share_of_step"], 4) for n, v in prof.items() if not n.startswith("_")}
        # the launch-bound YAML batch, for the record
        cfg_batch = 128
        tr_small, _ = make_trainer(cfg_batch, dev, layer_seed=rank, use_amp=(args.precision == "bf16"))
        small_rows = synth(16 * cfg_batch, D_MODEL, seed=4321 + rank).to(dev)
        small = [small_rows[i * cfg_batch:(i + 1) * cfg_batch].contiguous() for i in range(16)]
        sec_small, _ = time_steps(tr_small, small, 100, 10, False)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
            "scaling": "strong" if wl["dp"] else "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": args.batch * D_MODEL * 4 + 48,   # batch + control block
                    "d2h_bytes_per_step": 32, "ms_per_step": 1e3 * sec_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "roofline": roof,
            "kernel_shares": shares,
            "kernel_avg_ms": {n: round(v["avg_ms"], 4) for n, v in prof.items() if not n.startswith("_")},
            "yaml_batch": {"batch_rows": cfg_batch, "value": cfg_batch * 100 / sec_small, "unit": UNIT,
                           "ms_per_step": 1e3 * sec_small / 100},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_oracle_rate(args.batch, args.cpu_seconds, os.cpu_count() or 1)
    if dist_on:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
