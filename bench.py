#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: Mpix/s of GaussianBlur 5x5 on 3840x2160 BGR u8.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl reference]

A "step" is one pass of the hot path over one batch of F distinct 4K frames per GPU
(weak scaling: every rank owns F frames, no data-path collective; the only collective
is one broadcast of the filter taps before the first launch).

  value      whole-job Mpix/s, frames resident in HBM, ONE kernel launch per step,
             CUDA events on the library's stream, max over ranks.
  e2e        the same metric through the C-ABI batch call with HOST (pinned) Mats:
             H2D + kernel + D2H inside the timed region.
  roofline   6 algorithmic bytes/pixel (3 read + 3 written) / measured launch time
             against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  the CPU oracle (a C port of the op's definition; the reference's Rust
             has no GaussianBlur and no toolchain here) on a bounded sample, rank 0, N=1.

`--impl reference` times the CPU arm alone with the same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS, CN = 2160, 3840, 3
PIX = ROWS * COLS
ALGO_BYTES_PER_PIXEL = 6  # 3 B read + 3 B written
METRIC = "Mpix/s GaussianBlur 5x5 4K BGR u8"
TAPS_Q8 = [16, 64, 96, 64, 16]  # cv::getGaussianKernel(5, 0) * 256


# ---- multi-rank host logic (exercised on gloo by tests/test_dist_gloo.py) ---------------
def shard_frames(n_frames: int, rank: int, world: int) -> list[int]:
    """frame j -> rank j mod N."""
    return list(range(rank, n_frames, world))


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


def broadcast_taps(taps, device="cuda"):
    """The one collective of the path: rank 0's filter taps to every rank."""
    import torch

    d = _dist()
    t = torch.zeros(len(TAPS_Q8), dtype=torch.int32, device=device)
    if taps is not None:
        t.copy_(torch.as_tensor(np.asarray(taps, dtype=np.int32)))
    if d is not None:
        d.broadcast(t, src=0)
    return t.cpu().numpy()


def reduce_max(x: float, device="cuda") -> float:
    import torch

    d = _dist()
    t = torch.tensor([x], dtype=torch.float64, device=device)
    if d is not None:
        d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(x: float, device="cuda") -> float:
    import torch

    d = _dist()
    t = torch.tensor([x], dtype=torch.float64, device=device)
    if d is not None:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return float(t.item())


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()


# ---- clocks -----------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU arm ---------------------------------------------------------------------------------------
class CpuArm:
    """The oracle's GaussianBlur 5x5 on one SplitMix64 4K frame, row-parallel over `threads`.
    tuned=True times the auto-vectorised, bit-identical restatement instead of the definition port."""

    def __init__(self, threads: int, seed: int = 2, tuned: bool = False):
        from oracle import pyoracle as O

        O.build()
        self.O, self.threads, self.tuned = O, threads, tuned
        self.img = O.fill_u8(seed, PIX * CN).reshape(ROWS, COLS, CN)

    def run(self, frames: int) -> float:
        """Filters `frames` frames; returns seconds."""
        self.O.set_threads(self.threads)
        t0 = time.perf_counter()
        for _ in range(frames):
            if self.tuned:
                self.O.gaussian5_fast(self.img)
            else:
                self.O.gaussian_blur(self.img, (5, 5))
        dt = time.perf_counter() - t0
        self.O.set_threads(1)
        return dt


def cpu_gaussian_mpix(frames: int, threads: int, min_seconds: float = 0.0, tuned: bool = False) -> tuple[float, float, int]:
    """(Mpix/s, seconds, frames filtered): `frames` at a time until `min_seconds` of CPU work are on the clock."""
    arm = CpuArm(threads, tuned=tuned)
    arm.run(1)  # warm-up: page faults, thread start
    dt, done = 0.0, 0
    while done == 0 or dt < min_seconds:
        dt += arm.run(frames)
        done += frames
    return done * PIX / dt / 1e6, dt, done


def pcie_copy_peak(dev, nbytes: int, reps: int = 6) -> dict:
    """What the link itself gives: plain pinned<->device copies of one e2e step's bytes, each direction alone and
    both at once on two streams (CUDA events, best of `reps`).  The e2e figure is judged against `duplex_gbs_each`."""
    import torch

    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(up: bool, down: bool) -> float:
        best = 1e30
        for _ in range(reps):
            torch.cuda.synchronize(dev)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            s1.wait_event(e0)
            s2.wait_event(e0)
            if up:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
            e1.record(s1)
            e2.record(s2)
            torch.cuda.synchronize(dev)
            best = min(best, max(e0.elapsed_time(e1), e0.elapsed_time(e2)))
        return nbytes / (best * 1e-3) / 1e9

    return {"h2d_gbs": timed(True, False), "d2h_gbs": timed(False, True), "duplex_gbs_each": timed(True, True),
            "bytes_each_way": nbytes}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args) -> None:
    """`--impl reference`: the CPU implementation of the path on the host cores.  RustCV
    has no GaussianBlur and its Rust cannot be built here (no rustc), so this is the
    oracle port (kind "port"), row-parallel over all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    frames_per_step = 4
    arm = CpuArm(cores)
    for _ in range(max(args.warmup, 1)):
        arm.run(1)
    secs = 0.0
    for _ in range(args.steps):
        secs += arm.run(frames_per_step)
    mpix = args.steps * frames_per_step * PIX / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "GaussianBlur 5x5 sigma=0 REFLECT_101, 3840x2160 BGR u8 (BASELINE.json configs[1])",
                   "frames_per_step": frames_per_step},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {frames_per_step} SplitMix64 4K frames, oracle C port "
                                   f"(RustCV has no GaussianBlur; no rustc here), {cores} row-parallel threads"},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local: int) -> str:
    """Pins this rank to the CPUs local to its GPU (sysfs local_cpulist) BEFORE any pinned host
    memory is allocated, so the staging buffers are first-touched on the GPU's own NUMA node."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    except Exception:
        try:
            bus = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                                          text=True).strip()
        except Exception as e:  # noqa: BLE001
            return f"unbound ({e})"
    try:
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvidia-smi prints an 8-digit domain
            bus = bus[4:]
        path = f"/sys/bus/pci/devices/{bus}/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"bound to {len(allowed)} CPUs local to {bus}"
        return "unbound (no overlap with the allowed CPU set)"
    except Exception as e:  # noqa: BLE001
        return f"unbound ({e})"


# ---- GPU arm ---------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=32, help="4K frames per GPU per step")
    ap.add_argument("--e2e-frames", type=int, default=32, help="host frames per GPU per e2e step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: rustcv_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import ctypes as C

    import rustcv_b200 as R
    from oracle import pyoracle as O  # input generator (SplitMix64) + cpu_baseline leg only
    from rustcv_b200 import _ffi as F

    R.imgproc.init(local)
    dev = torch.device("cuda", local)

    # the path's one collective: filter taps from rank 0
    taps = broadcast_taps(TAPS_Q8 if rank == 0 else None, device=dev)
    assert taps.tolist() == TAPS_Q8

    # ---- data: F distinct frames per rank, frame j of the job uses seed 2 + j ----------
    F_ = args.frames
    src = R.Mat.device_batch(F_, ROWS, COLS, CN)
    dst = R.Mat.device_batch(F_, ROWS, COLS, CN)
    host = R.Mat.pinned(ROWS, COLS, CN)
    my_frames = shard_frames(F_ * world, rank, world)
    for i, j in enumerate(my_frames):
        host.data[:] = O.fill_u8(2 + j, PIX * CN)
        F.check(F.lib.rcv_mat_upload(C.byref(host.c()), C.byref(src[i].c())))

    stream = torch.cuda.ExternalStream(R.imgproc.stream_ptr(local), device=dev)
    R.imgproc.set_blocking(False)

    def step():
        R.imgproc.gaussian_blur_batch(src, dst, (5, 5), 0.0, 0.0)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    R.imgproc.sync(local)
    torch.cuda.synchronize()
    barrier()

    launches0 = R.imgproc.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    R.imgproc.sync(local)
    torch.cuda.synchronize()
    launches = R.imgproc.launch_count() - launches0
    ms_local = e0.elapsed_time(e1)
    barrier()
    ms = reduce_max(ms_local, device=dev)
    ms_per_step = ms / args.steps
    total_pix = reduce_sum(float(F_ * PIX), device=dev)  # pixels per step over all ranks
    value = total_pix / (ms_per_step * 1e-3) / 1e6

    # sanity: frame 0 of rank 0 reproduces the golden CRC (seed 2)
    crc_ok = None
    if rank == 0:
        crc_ok = O.crc32(dst[0].to_numpy()) == 0x827081C8

    # ---- e2e: host (pinned) Mats through the public batch call -------------------------
    R.imgproc.set_blocking(True)
    E = args.e2e_frames
    hsrc = [R.Mat.pinned(ROWS, COLS, CN) for _ in range(E)]
    hdst = [R.Mat.pinned(ROWS, COLS, CN) for _ in range(E)]
    for i in range(E):
        hsrc[i].data[:] = O.fill_u8(2 + my_frames[i % len(my_frames)], PIX * CN)

    def e2e_step():
        R.imgproc.gaussian_blur_batch(hsrc, hdst, (5, 5), 0.0, 0.0)

    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s_local = time.perf_counter() - t0
    barrier()
    e2e_s = reduce_max(e2e_s_local, device=dev)
    e2e_pix = reduce_sum(float(E * PIX), device=dev)
    e2e_value = e2e_pix * args.steps / e2e_s / 1e6
    # the reference API's own shape: ONE synchronous call per frame on a pinned host Mat (banded pipeline)
    for _ in range(3):
        R.imgproc.gaussian_blur(hsrc[0], hdst[0], (5, 5), 0.0)
    t0 = time.perf_counter()
    for _ in range(20):
        R.imgproc.gaussian_blur(hsrc[0], hdst[0], (5, 5), 0.0)
    single_ms = (time.perf_counter() - t0) / 20 * 1e3
    clocks = sampler.stop() if rank == 0 else None  # sampled from warm-up through both timed regions
    pcie = pcie_copy_peak(dev, E * PIX * CN) if rank == 0 else None  # after the timed regions, outside them
    e2e_ok = None
    if rank == 0:
        e2e_ok = O.crc32(hdst[0].to_numpy()) == 0x827081C8

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # one launch per step on every rank: per-launch time = ms_per_step
        achieved = ALGO_BYTES_PER_PIXEL * F_ * PIX / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("frames_per_launch") == F_:
                traffic = tj.get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "GaussianBlur 5x5 sigma=0 REFLECT_101, 3840x2160 BGR u8 (BASELINE.json configs[1])",
                       "frames_per_gpu_per_step": F_, "launches_per_step_per_gpu": 1,
                       "l2": f"inputs larger than L2: {2 * F_ * PIX * CN / 1e6:.0f} MB touched per step vs 126 MB L2",
                       "input": "SplitMix64 seed 2+j per frame", "parallelism": f"frames sharded, {world} rank(s)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "k_strip<Gauss5Op<3>>",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PIXEL * F_ * PIX},
            "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": E * PIX * CN * world,
                    "d2h_bytes_per_step": E * PIX * CN * world, "frames_per_gpu_per_step": E,
                    "api": "rcv_gaussian_blur_batch on pinned host Mats", "host_binding": numa,
                    "single_frame_call_ms": single_ms,
                    "link": pcie,
                    "achieved_gbs_each_way": e2e_value * 1e6 * CN / 1e9 / world,
                    "frac_of_duplex_copy": (e2e_value * 1e6 * CN / 1e9 / world) / pcie["duplex_gbs_each"]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": {"device_frame0_crc_827081c8": crc_ok, "e2e_frame0_crc_827081c8": e2e_ok},
        }
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)  # the CPU arm gets every host core again
            cores = host_cores()
            v_all, s_all, n_all = cpu_gaussian_mpix(16 if cores >= 8 else 4, cores, min_seconds=8.0)
            v_one, s_one, n_one = cpu_gaussian_mpix(4, 1, min_seconds=4.0)
            t_all, ts_all, tn_all = cpu_gaussian_mpix(16, cores, min_seconds=3.0, tuned=True)
            t_one, ts_one, tn_one = cpu_gaussian_mpix(8, 1, min_seconds=2.0, tuned=True)
            line["cpu_baseline"] = {"value": v_all, "unit": "Mpix/s", "cores": cores, "kind": "port",
                                    "tuned_port": {"value": t_all, "value_1_thread": t_one, "cores": cores,
                                                   "what": "the same result from an auto-vectorised restatement "
                                                           "(u16 vertical pass, branch-free horizontal pass; "
                                                           f"bit-identical, gcc -O3, no -march): {tn_all} frames in "
                                                           f"{ts_all:.1f} s; not the definition, reported so the CPU "
                                                           "figure is not an artefact of scalar code"},
                                    "sample": f"oracle C port, {cores} threads x {n_all} frames of the same workload "
                                              f"({s_all:.1f} s); 1 thread x {n_one} frames = {v_one:.0f} Mpix/s ({s_one:.1f} s)",
                                    "value_1_thread": v_one}
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
