#!/usr/bin/env python3
"""bench.py -- measures the colour-conversion hot path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c4|c5] [--impl b200|reference]

A "step" is one pass of the hot path over one synthetic frame (or batch) of the workload.  Default workload c2 is
the configuration BASELINE.json quotes its metric on: 7680x4320 RGB32f -> 12-bit Rec.2100 PQ YCbCr 4:2:0.

  value      whole-job Gpixels/s with inputs already resident in HBM, K steps between two CUDA events on the
             launching stream, barrier + synchronize on both sides, max over ranks.
  e2e        the same metric through the host-pointer C-ABI call the plug-in binds (avifgpu_encode_rows /
             avifgpu_decode_rows): pinned host rows -> PCIe -> kernel -> PCIe -> pinned host planes, every step.
  roofline   algorithmic bytes per launch / mean launch duration (CUDA events around every launch) against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference's CPU loop (oracle/_ref where the reference has the stage, the C restatement for
             the libheif stage) on this box's host cores, on a bounded sample -- reported, not a target.

With --impl reference only the CPU arm runs (rank 0 under torchrun; other ranks exit 0).
Multi-GPU (torchrun, one rank per GPU): frames / tiles are independent, so ranks share nothing on the data path
("scaling": "weak": every rank converts its own frame; --mode tile splits ONE frame into row blocks and gathers
the planes over NCCL, reported separately in `gather`).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "avif-format_b200", "python"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

from avifgpu import abi  # noqa: E402

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent


# ---- workloads ------------------------------------------------------------------------------------------------------

def nclx_2020(transfer):
    return abi.Nclx(1, abi.PRIMARIES_BT2020, transfer, abi.MATRIX_BT2020_NCL, 1)


class Workload:
    """Describes one BASELINE.json configuration: geometry, descriptions, byte counts, synthetic data."""

    def __init__(self, key):
        self.key = key
        if key == "c2":
            self.name = "7680x4320 RGB32f -> 12-bit Rec.2100 PQ YCbCr 4:2:0 (BT.2020 NCL, full range, box down-filter, peak 80 nit)"
            self.direction, self.w, self.h, self.batch = "encode", 7680, 4320, 1
            self.enc = abi.EncodeDesc(self.w, self.h, 32, 3, abi.ALPHA_NONE, 12, abi.TRANSFER_PQ, 80, abi.LAYOUT_PLANAR_YCBCR,
                                      abi.CHROMA_420, abi.DOWN_FILTER_BOX, abi.GRAY16_LUT, nclx_2020(abi.TRANSFER_CHAR_PQ))
            self.bytes_per_pixel = 12 + 3  # in 12 B/px, out Y 2 + CbCr 2*2/4
            self.dtype = "f32"
        elif key == "c3":
            self.name = "7680x4320 10-bit HLG YCbCr 4:2:0 -> RGB32f (BT.2020 NCL, full range, OOTF gamma 1.2 @ 1000 nit)"
            self.direction, self.w, self.h, self.batch = "decode", 7680, 4320, 1
            self.dec = abi.DecodeDesc(self.w, self.h, abi.COLORSPACE_YCBCR, abi.CHROMA_420, 10, abi.ALPHA_NONE, 32,
                                      nclx_2020(abi.TRANSFER_CHAR_HLG), 1, 1.2, 1000, 80)
            self.bytes_per_pixel = 3 + 12
            self.dtype = "f32"
        elif key == "c4":
            self.name = "16384x16384 RGBA16 -> 10-bit YCbCr 4:2:2 + alpha plane (BT.601, straight alpha)"
            self.direction, self.w, self.h, self.batch = "encode", 16384, 16384, 1
            self.enc = abi.EncodeDesc(self.w, self.h, 16, 4, abi.ALPHA_STRAIGHT, 10, abi.TRANSFER_CLIP, 80, abi.LAYOUT_PLANAR_YCBCR,
                                      abi.CHROMA_422, abi.DOWN_FILTER_BOX, abi.GRAY16_LUT, None)
            self.bytes_per_pixel = 8 + 6
            self.dtype = "f32"
        elif key == "c5":
            self.name = "batch of 4096x4096 Gray16 -> 12-bit monochrome SMPTE 428-1 (32 images per GPU)"
            self.direction, self.w, self.h, self.batch = "encode", 4096, 4096, 32
            self.enc = abi.EncodeDesc(self.w, self.h * self.batch, 16, 1, abi.ALPHA_NONE, 12, gray16_curve=abi.GRAY16_SMPTE428)
            self.bytes_per_pixel = 2 + 2
            self.dtype = "f32"
        else:
            raise SystemExit(f"unknown workload {key}")
        self.rows_total = self.h * self.batch
        self.pixels = self.w * self.rows_total
        self.algorithmic_bytes = self.pixels * self.bytes_per_pixel

    # -- descriptions for a row block presented as an image (CPU sample, multi-GPU tiles)
    def encode_desc(self, rows):
        return self.enc.copy(height=rows)

    def decode_desc(self, rows):
        return self.dec.copy(height=rows)

    # -- synthetic data, generated on the device with a seeded generator
    def make_device_input(self, torch, device, seed):
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        if self.key == "c2" and getattr(self, "smooth", False):
            # A smooth frame (supplementary, --data smooth): horizontal ramps per channel with 0.1 % noise, so that
            # neighbouring pixels fall into the same step-table buckets the way natural images do.
            x = torch.linspace(0.0, 1.0, self.w, device=device).repeat_interleave(3)
            phase = torch.tensor([0.05, 0.35, 0.65], device=device).repeat(self.w)
            base = 0.02 + 0.9 * torch.remainder(x + phase, 1.0)
            noise = 1.0 + 1.0e-3 * (torch.rand((self.h, self.w * 3), generator=g, device=device) - 0.5)
            return [(base[None, :] * noise).contiguous()]
        if self.key == "c2":
            n = (self.h, self.w * 3)
            kind = torch.rand(n, generator=g, device=device)
            v = torch.rand(n, generator=g, device=device)
            lo, hi = float(np.log(1e-4)), float(np.log(4.0))
            logu = torch.exp(torch.rand(n, generator=g, device=device) * (hi - lo) + lo)
            v = torch.where(kind < 0.20, logu, v)
            v = torch.where((kind >= 0.20) & (kind < 0.25), -torch.rand(n, generator=g, device=device), v)
            v = torch.where((kind >= 0.25) & (kind < 0.30), torch.round(torch.rand(n, generator=g, device=device)), v)
            return [v.contiguous()]
        if self.key == "c3":
            cw, ch = (self.w + 1) // 2, (self.h + 1) // 2
            y = torch.randint(0, 1024, (self.h, self.w), generator=g, device=device, dtype=torch.int16)
            cb = torch.randint(0, 1024, (ch, cw), generator=g, device=device, dtype=torch.int16)
            cr = torch.randint(0, 1024, (ch, cw), generator=g, device=device, dtype=torch.int16)
            return [y, cb, cr]
        if self.key == "c4":
            v = torch.randint(0, 32769, (self.h, self.w * 4), generator=g, device=device, dtype=torch.int32).to(torch.int16)
            return [v]
        v = torch.randint(0, 32769, (self.rows_total, self.w), generator=g, device=device, dtype=torch.int32).to(torch.int16)
        return [v]

    def make_device_output(self, torch, device):
        if self.direction == "decode":
            return [torch.empty((self.h, self.w * 3), dtype=torch.float32, device=device)]
        shapes = abi.encode_plane_shapes(self.enc)
        dt = torch.int16 if self.enc.image_bit_depth > 8 else torch.uint8
        return [None if s is None else torch.empty(s, dtype=dt, device=device) for s in shapes]


# ---- helpers ----------------------------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md 'clocks line'), polled through NVML
    from a thread (nvidia-smi -lms cannot resolve a region of a few milliseconds); nvidia-smi is the fallback."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.running = False
        self.nvml = None
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            physical = self.index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    physical = int(ids[self.index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(physical)
            self.nvml = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None
            return
        self.running = True
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        n = self.nvml
        while self.running:
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(0.0005)

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            return float(out[0]), float(out[1])
        except Exception:
            return None, None

    def stop(self, t0, t1):
        if self.nvml is None:
            sm, smax = self._smi_once()
            return {"sm_mhz": sm, "sm_max_mhz": smax, "reasons": ["nvml unavailable: one nvidia-smi sample taken after the timed region"], "samples": 0}
        self.running = False
        self.thread.join(timeout=1.0)
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        if not inside:
            inside = self.samples[-1:]
        mask = 0
        for s in inside:
            mask |= s[2]
        return {"sm_mhz": statistics.median([s[1] for s in inside]) if inside else None, "sm_max_mhz": self.max_mhz,
                "reasons": [name for name, bit in self.REASONS if mask & bit], "samples": len(inside)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(workload_key):
    """dram bytes per launch from the committed ncu capture of the same command, if one has been recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload_key)
    except Exception:
        return None


# ---- the CPU arm (reference's CPU loop; the only place bench.py touches oracle/) -------------------------------------------

class CpuArm:
    def __init__(self, workload):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle  # noqa: E402  (checker / CPU baseline only)
        self.oracle = oracle
        self.port = oracle.load_restatement()
        self.ref = oracle.load_reference()
        self.w = workload
        self.threads = os.cpu_count() or 1

    def kind(self):
        return "reference" if self.ref is not None else "port"

    def describe(self, rows):
        wl = self.w
        if wl.key == "c2":
            stage = ("CreateHeifImageRGBThirtyTwoBit via the compiled reference TUs, then the libheif stage (matrix + 4:2:0 box; not in "
                     "the reference tree) via the C restatement") if self.ref is not None else "C restatement (fused)"
        elif wl.key == "c3":
            stage = "ReadHeifImageRGBThirtyTwoBit via the compiled reference TUs" if self.ref is not None else "C restatement"
        elif wl.key == "c4":
            stage = ("CreateHeifImageRGBSixteenBit via the compiled reference TUs + C restatement of the libheif stage"
                     if self.ref is not None else "C restatement (fused)")
        else:
            stage = "C restatement (Gray16 -> SMPTE 428 is this project's composition; the reference has no such path)"
        return f"{rows} rows x {wl.w} px of the workload ({rows * wl.w / 1e6:.2f} Mpx), {self.threads} threads, row-block parallel; {stage}"

    def _input(self, rows, seed):
        rng = np.random.default_rng(seed)
        wl = self.w
        if wl.key == "c2":
            import cases
            return cases.float_host_rows(rng, rows, wl.w, 3)
        if wl.key == "c3":
            cw, ch = (wl.w + 1) // 2, (rows + 1) // 2
            return [rng.integers(0, 1024, (rows, wl.w)).astype(np.uint16), rng.integers(0, 1024, (ch, cw)).astype(np.uint16),
                    rng.integers(0, 1024, (ch, cw)).astype(np.uint16), None]
        if wl.key == "c4":
            return rng.integers(0, 32769, (rows, wl.w * 4)).astype(np.uint16)
        return rng.integers(0, 32769, (rows, wl.w)).astype(np.uint16)

    def run_once(self, rows, data):
        """Converts `rows` rows; returns seconds."""
        wl, t = self.w, self.threads
        t0 = time.perf_counter()
        if wl.direction == "decode":
            (self.ref or self.port).decode(wl.decode_desc(rows), data, threads=t)
        elif wl.key in ("c2", "c4") and self.ref is not None:
            desc = wl.encode_desc(rows)
            inter = self.ref.encode(desc.copy(layout=abi.LAYOUT_REFERENCE), data, threads=t)[0]
            self.port.rgb_codes_to_ycbcr(desc, inter, threads=t)
        else:
            self.port.encode(wl.encode_desc(rows), data, threads=t)
        return time.perf_counter() - t0

    def sample(self, target_seconds, seed=99):
        """Sizes a row block so one conversion takes about target_seconds; returns (rows, seconds)."""
        wl = self.w
        probe_rows = min(wl.rows_total, max(2 * self.threads, 16)) & ~1
        probe_rows = max(probe_rows, 2)
        data = self._input(probe_rows, seed)
        self.run_once(probe_rows, data)
        t_probe = max(self.run_once(probe_rows, data), 1e-4)
        rows = int(probe_rows * target_seconds / t_probe) & ~1
        rows = max(min(rows, wl.rows_total), probe_rows)
        data = self._input(rows, seed + 1)
        return rows, data


def run_reference_impl(args, workload, rank, world):
    if rank != 0:
        return
    arm = CpuArm(workload)
    budget = 150.0  # seconds for the whole run
    per_step = min(max(budget / (args.steps + args.warmup + 1), 0.5), 15.0)
    rows, data = arm.sample(per_step)
    for _ in range(args.warmup):
        arm.run_once(rows, data)
    t = 0.0
    for _ in range(args.steps):
        t += arm.run_once(rows, data)
    value = rows * workload.w * args.steps / t / 1e9
    line = {
        "impl": "reference", "metric": "Gpixels/s", "value": value, "unit": "Gpx/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": workload.dtype, "data": "synthetic",
        "config": {"workload": workload.name, "sample_rows_per_step": rows, "host_threads": arm.threads},
        "cpu_baseline": {"value": value, "unit": "Gpx/s", "cores": arm.threads, "kind": arm.kind(), "sample": arm.describe(rows),
                         "libm": arm.port.libm_version()},
        "e2e": {"value": value, "unit": "Gpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- the GPU arm --------------------------------------------------------------------------------------------------------

def run_b200(args, workload, rank, world, local_rank):
    import torch
    import avifgpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the avifgpu path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=device)

    gpu = avifgpu.Context(local_rank)
    wl = workload
    if wl.direction == "encode":
        gpu.prepare_encode(wl.enc)  # a frame pipeline builds its step tables up front (one-off, ~40 ms)
    copies = 3  # rotate distinct frames so nothing of a previous step survives in the 126 MB L2
    seed0 = {"c2": 2, "c3": 3, "c4": 4, "c5": 5}[wl.key] * 1000 + 1234
    inputs = [wl.make_device_input(torch, device, seed0 + 17 * i + 101 * rank) for i in range(copies)]
    outputs = [wl.make_device_output(torch, device) for _ in range(copies)]
    stream = torch.cuda.current_stream(device)
    stream_handle = stream.cuda_stream

    def launch(i):
        src, dst = inputs[i % copies], outputs[i % copies]
        if wl.direction == "encode":
            rows = src[0]
            gpu.encode_device(wl.enc, rows.data_ptr(), rows.stride(0) * rows.element_size(), avifgpu.planes_from_tensors(dst),
                              stream=stream_handle)
        else:
            planes = avifgpu.planes_from_tensors([src[0], src[1], src[2], None])
            out = dst[0]
            gpu.decode_device(wl.dec, planes, out.data_ptr(), out.stride(0) * out.element_size(), stream=stream_handle)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    # -- warm-up
    for i in range(max(args.warmup, 3)):
        launch(i)
    barrier()

    # -- timed region: K steps, device-resident inputs
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    events = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches_before = gpu.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    events[0].record(stream)
    for i in range(args.steps):
        launch(i)
        events[i + 1].record(stream)
    torch.cuda.synchronize(device)
    t_wall1 = time.perf_counter()
    barrier()
    launches = gpu.launch_count() - launches_before
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    elapsed_ms = events[0].elapsed_time(events[-1])
    per_launch_ms = [events[i].elapsed_time(events[i + 1]) for i in range(args.steps)]
    if dist is not None:
        t = torch.tensor([elapsed_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * wl.pixels * args.steps / (elapsed_ms * 1e-3) / 1e9

    # -- e2e: the host-pointer C-ABI call, pinned host buffers, copies inside the timed region
    e2e_steps = max(1, min(args.steps, 10))
    if wl.direction == "encode":
        host_in = torch.empty(inputs[0][0].shape, dtype=inputs[0][0].dtype, pin_memory=True)
        host_in.copy_(inputs[0][0])
        shapes = abi.encode_plane_shapes(wl.enc)
        dt = torch.int16 if wl.enc.image_bit_depth > 8 else torch.uint8
        host_out = [None if s is None else torch.empty(s, dtype=dt, pin_memory=True) for s in shapes]
        host_planes = abi.Planes()
        for k, t in enumerate(host_out):
            if t is not None:
                host_planes.data[k] = t.data_ptr()
                host_planes.stride[k] = t.stride(0) * t.element_size()
        h2d = host_in.numel() * host_in.element_size()
        d2h = sum(t.numel() * t.element_size() for t in host_out if t is not None)

        def e2e_step():
            gpu._check(gpu.lib.avifgpu_encode_rows(gpu.handle, C.byref(wl.enc), host_in.data_ptr(), host_in.stride(0) * host_in.element_size(),
                                                   0, wl.rows_total, C.byref(host_planes)))
    else:
        host_src = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in inputs[0]]
        for hs, t in zip(host_src, inputs[0]):
            hs.copy_(t)
        host_rows = torch.empty(outputs[0][0].shape, dtype=torch.float32, pin_memory=True)
        host_planes = abi.Planes()
        for k, t in enumerate(host_src):
            host_planes.data[k] = t.data_ptr()
            host_planes.stride[k] = t.stride(0) * t.element_size()
        h2d = sum(t.numel() * t.element_size() for t in host_src)
        d2h = host_rows.numel() * 4

        def e2e_step():
            gpu._check(gpu.lib.avifgpu_decode_rows(gpu.handle, C.byref(wl.dec), C.byref(host_planes), 0, wl.h, host_rows.data_ptr(),
                                                   host_rows.stride(0) * 4))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(device)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * wl.pixels * e2e_steps / e2e_s / 1e9

    # -- optional: one frame split in row blocks + NCCL gather of the planes to every rank
    gather = None
    if dist is not None and args.mode == "tile" and wl.direction == "encode":
        gather = run_tile_mode(torch, dist, gpu, avifgpu, wl, device, rank, world, args, stream_handle)

    if rank == 0:
        peak, peak_source = measured_peak()
        mean_launch_ms = statistics.mean(per_launch_ms)
        achieved = wl.algorithmic_bytes / (mean_launch_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            arm = CpuArm(wl)
            rows, data = arm.sample(12.0)
            seconds = arm.run_once(rows, data)
            cpu = {"value": rows * wl.w / seconds / 1e9, "unit": "Gpx/s", "cores": arm.threads, "kind": arm.kind(),
                   "sample": arm.describe(rows), "seconds": seconds, "libm": arm.port.libm_version()}
        line = {
            "metric": "Gpixels/s", "value": value, "unit": "Gpx/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": wl.dtype, "data": "synthetic" if not getattr(wl, "smooth", False) else "synthetic (smooth ramps + 0.1 % noise; supplementary)",
            "config": {"workload": wl.name, "pixels_per_step_per_gpu": wl.pixels, "parallelism": f"{world} independent frame(s), one per GPU",
                       "l2": f"{copies} rotating input/output sets of {wl.algorithmic_bytes / 1e6:.0f} MB each (> 126 MB L2)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(wl.key), "peak_source": peak_source,
                         "algorithmic_bytes_per_launch": wl.algorithmic_bytes, "mean_launch_ms": mean_launch_ms,
                         "min_launch_ms": min(per_launch_ms)},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Gpx/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if gather is not None:
            line["gather"] = gather
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    gpu.close()


def run_tile_mode(torch, dist, gpu, avifgpu, wl, device, rank, world, args, stream_handle):
    """ONE frame, rank r converts row block r (avifgpu.sharding.row_blocks: even boundaries, no halo), then the planar
    buffer is assembled on every rank with one all_gather per plane.  Returns timings (convert-only and
    convert+gather) for rank 0 to print, after checking the gathered planes against a single-GPU conversion."""
    from avifgpu import sharding
    blocks = sharding.row_blocks(wl.rows_total, world)
    y0, n = blocks[rank]
    desc = sharding.block_desc(wl.enc, n)
    full = wl.make_device_input(torch, device, 4242)[0]  # same seed on every rank: the same frame
    block = full[y0:y0 + n]
    shapes = sharding.max_block_plane_shapes(wl.enc, blocks)
    dt = torch.int16 if wl.enc.image_bit_depth > 8 else torch.uint8
    local = [None if s is None else torch.zeros(s, dtype=dt, device=device) for s in shapes]

    def convert():
        if n > 0:
            gpu.encode_device(desc, block.data_ptr(), block.stride(0) * block.element_size(), avifgpu.planes_from_tensors(local),
                              stream=stream_handle)

    def gather():
        return sharding.gather_encode_planes(dist, torch, wl.enc, blocks, local)

    convert()
    planes = gather()
    torch.cuda.synchronize(device)
    # correctness of the tiling + gather: compare with the whole frame converted on this GPU alone
    whole = wl.make_device_output(torch, device)
    gpu.encode_device(wl.enc, full.data_ptr(), full.stride(0) * full.element_size(), avifgpu.planes_from_tensors(whole), stream=stream_handle)
    torch.cuda.synchronize(device)
    identical = all(torch.equal(a, b) for a, b in zip(planes, whole) if a is not None)
    del planes, whole
    for _ in range(2):
        convert()
        gather()
    dist.barrier()
    torch.cuda.synchronize(device)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(args.steps):
        convert()
    e[1].record()
    for _ in range(args.steps):
        convert()
        gather()
    e[2].record()
    torch.cuda.synchronize(device)
    t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    convert_ms, both_ms = (float(v) / args.steps for v in t.tolist())
    report = {"mode": "one frame, even row-block tiles", "identical_to_single_gpu": bool(identical),
              "convert_ms": convert_ms, "convert_plus_all_gather_ms": both_ms,
              "convert_gpx_s": wl.pixels / (convert_ms * 1e-3) / 1e9, "convert_plus_all_gather_gpx_s": wl.pixels / (both_ms * 1e-3) / 1e9}

    # Fused placement: the conversion kernels store straight into rank 0's planes over NVLink (CUDA IPC peer mapping),
    # so the assembled image exists on the owner when the kernels end -- no gather pass.
    try:
        peer = sharding.PeerPlanes(dist, wl.enc, rank, owner=0)
    except Exception as error:  # no peer access on this box: report it, keep the all_gather numbers
        report["peer_placement"] = f"unavailable: {error}"
        return report
    peer_planes = peer.planes()
    block_ptr, block_stride = block.data_ptr(), block.stride(0) * block.element_size()

    def convert_place():
        if n > 0:
            gpu.encode_device(wl.enc, block_ptr, block_stride, peer_planes, y0=y0, nrows=n, stream=stream_handle)

    convert_place()
    torch.cuda.synchronize(device)
    dist.barrier()
    placed_identical = None
    if rank == 0:
        whole = wl.make_device_output(torch, device)
        gpu.encode_device(wl.enc, full.data_ptr(), full.stride(0) * full.element_size(), avifgpu.planes_from_tensors(whole), stream=stream_handle)
        torch.cuda.synchronize(device)
        placed = peer.owner_tensors(torch, device)
        placed_identical = all(torch.equal(a, b) for a, b in zip(placed, whole) if a is not None)
        del placed, whole
    for _ in range(2):
        convert_place()
    torch.cuda.synchronize(device)
    dist.barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(args.steps):
        convert_place()
    e[1].record()
    torch.cuda.synchronize(device)
    t = torch.tensor([e[0].elapsed_time(e[1])], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    place_ms = float(t.item()) / args.steps
    dist.barrier()
    peer.close()
    report.update({"peer_placement": "kernels store into rank 0's planes over NVLink (CUDA IPC); the image is assembled when they end",
                   "placed_identical_to_single_gpu": placed_identical, "convert_and_place_ms": place_ms,
                   "convert_and_place_gpx_s": wl.pixels / (place_ms * 1e-3) / 1e9})
    return report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--workload", choices=("c2", "c3", "c4", "c5"), default="c2")
    ap.add_argument("--mode", choices=("frames", "tile"), default="frames")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--data", choices=("random", "smooth"), default="random",
                    help="c2 only: 'random' (default, SURVEY 8(d): the worst case for the table look-ups) or 'smooth' (supplementary)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = Workload(args.workload)
    workload.smooth = args.data == "smooth"
    if args.impl == "reference":
        run_reference_impl(args, workload, rank, world)
        return
    run_b200(args, workload, rank, world, local_rank)


if __name__ == "__main__":
    main()
