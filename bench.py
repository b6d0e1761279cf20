#!/usr/bin/env python3
"""bench.py — headline benchmark: Mpath-samples/s of the spectral path-tracing hot path.

A "step" = one frame of BASELINE.json configs[1]: cornell-srgb 512x512, spp 64, hero-wavelength (4 λ),
RENDER_MODE_SPECTRAL_OURS, CIE 1931, explicit light sampling, MAX_DEPTH 10, per-sample seeding.
  value : whole-job samples / time with scene, tables and texture already resident in HBM
          (trace kernel + in-order f64 accumulation + resolve to XYZA/sRGBA, results left on the device)
  e2e   : the same metric through the reference-facing C-ABI calls with HOST buffers — every step uploads
          the scene (incl. the 48 MiB RGB8 texture) and colour tables from pinned host memory and reads the
          XYZA + sRGBA framebuffers back (ssb_upload_scene + ssb_upload_color + ssb_render_frame)
Multi-GPU (torchrun, one rank per GPU): weak scaling — every rank renders its own 64 samples per pixel of the
same 512x512 frame (sample indices r*64..r*64+63), then ONE NCCL reduce(sum) of the f64 XYZA accumulators to
rank 0, which resolves the spp = 64*N image.
`--impl reference` times the reference's own multithreaded CPU renderer (oracle/_ref, built from /root/reference
by oracle/build_ref.py) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpath-samples/s"
JSON_OUT = None  # where the one JSON line goes (stdout unless main_ours redirected fd 1, see there)
SCENE, VARIANT = "cornell-srgb", "ours1931"
HEADLINE_SHA = "3ceea29f9217d925"  # f64 accumulators of the headline frame (seed 1) on one GPU: tools/ab.py prints it for every build
# SURVEY.md §8(d) algorithmic HBM bytes per path sample (wavefront model the north star names):
# 5.30 closest-hit stages x 192 B ray state read+write + 47 B texture sectors + 32 B f64 XYZA output
ALGO_BYTES_PER_SAMPLE = {"cornell-srgb": 5.30 * 192 + 47 + 32, "cornell": 5.30 * 192 + 32, "plane-srgb": 2.0 * 192 + 64 + 32}
JH_EXTRA_BYTES_PER_SAMPLE = {"cornell-srgb": 375.0, "cornell": 0.0, "plane-srgb": 512.0}  # 8 coefficient sectors per textured lookup


def source_sha():
    """sha256 over the device-code sources: ties an ncu-derived figure (profiles/trace_kernel_traffic.json) to the build
    it was captured on — bench.py refuses to quote DRAM traffic / instruction counts of another build."""
    import hashlib
    h = hashlib.sha256()
    src = os.path.join(ROOT, "simple-spectral_b200", "csrc")
    for f in sorted(os.listdir(src)):
        if f.endswith((".cu", ".cuh", ".hpp")):
            h.update(open(os.path.join(src, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "ssb200.h"), "rb").read())
    return h.hexdigest()[:16]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled every 5 ms from a thread (the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; nvidia-smi itself needs ~100 ms to
    start, longer than a short timed region), nvidia-smi -lms as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASON_BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.stop_flag, self.sm, self.bits, self.smax = None, False, [], 0, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                self.bits |= int(get(self.handle))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.thread.join(timeout=1)
            sm = sorted(self.sm)
            reasons = sorted(name for bit, name in self.REASON_BITS.items() if self.bits & bit)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "samples": len(sm), "reasons": reasons, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def data_root():
    host = importlib.import_module("simple-spectral_b200.host")
    return host.find_data_root()


# --------------------------------------------------------------------------------------------- reference arm
def ref_binary(variant=None):
    variant = variant or VARIANT
    p = os.path.join(ROOT, "oracle", "_ref", f"simple_spectral_{variant}")
    if not os.path.exists(p):
        raise SystemExit(f"{p} missing: run __graft_entry__.build() where /root/reference exists")
    return p


def run_reference_once(width, height, spp, threads=None, variant=None, scene=None):
    """Runs the UNMODIFIED reference renderer; returns its own 'Render completed in' time (excludes load)."""
    variant, scene = variant or VARIANT, scene or SCENE
    env = dict(os.environ)
    out = f"/tmp/ssb_ref_bench_{os.getpid()}.pfm"
    r = subprocess.run([ref_binary(variant), f"--scene={scene}", f"-w={width}", f"-h={height}", f"-spp={spp}", f"--output={out}"],
                       cwd=data_root(), env=env, capture_output=True, text=True)
    try:
        os.remove(out)
    except OSError:
        pass
    m = re.findall(r"Render completed in (?:(\d+) days \+ )?(\d+):(\d+):([0-9.]+)", r.stdout)
    if r.returncode != 0 or not m:
        raise SystemExit("reference run failed: " + r.stderr[-500:])
    d, h, mi, s = m[-1]
    return (int(d or 0) * 86400) + int(h) * 3600 + int(mi) * 60 + float(s)


def cpu_baseline(width, height, sample_spp):
    cores = os.cpu_count() or 1
    secs = run_reference_once(width, height, sample_spp)
    n = width * height * sample_spp
    return {"value": n / secs / 1e6, "unit": METRIC, "cores": cores, "kind": "reference",
            "sample": f"{SCENE} {width}x{height} spp{sample_spp} ({n} samples, {secs:.2f} s by the reference's own timer; "
                      f"unmodified reference sources, g++ -O3 -march=x86-64-v3, std::thread::hardware_concurrency()={cores} threads)"}


def main_reference(args, rank, world):
    if rank != 0:
        return 0
    spp = args.ref_spp
    times = []
    for i in range(args.warmup + args.steps):
        t = run_reference_once(args.width, args.height, spp)
        if i >= args.warmup:
            times.append(t)
    total = sum(times)
    n = args.width * args.height * spp
    value = n * len(times) / total / 1e6
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "reference scene (hard-coded geometry + shipped spectra/texture)",
        "config": {"workload": f"{SCENE} {args.width}x{args.height} hero-wavelength {VARIANT}; each step = one frame at spp{spp} "
                               f"(bounded sample of the spp{args.spp} workload)", "timer": "reference's own 'Render completed in' (excludes scene load)"},
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": cores, "kind": "reference",
                         "sample": f"{args.steps} frames of {SCENE} {args.width}x{args.height} spp{spp}, {cores} threads"},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------- our arm
def main_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    ssb = importlib.import_module("simple-spectral_b200")
    host = importlib.import_module("simple-spectral_b200.host")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # rank 0 prints ONE JSON line on stdout: NCCL writes its version banner (and anything NCCL_DEBUG asks for) to
        # fd 1 from native code, so fd 1 is pointed at stderr for the whole run and the JSON goes to the saved descriptor
        global JSON_OUT
        sys.stdout.flush()
        _json_fd = os.dup(1)
        os.dup2(2, 1)
        JSON_OUT = os.fdopen(_json_fd, "w")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, H, SPP = args.width, args.height, args.spp
    color = host.Color(None, *host.VARIANTS[VARIANT])
    scene = host.Scene(SCENE, color)
    ctx = ssb.Context(local_rank)
    ctx.upload_color(color.flat)
    ctx.upload_scene(scene.flat)
    # a dedicated (non-default) stream: the context issues every kernel/copy on it, and the timing events below are
    # recorded on the same stream (handle 0 = "legacy default stream" would mean "use the context's own stream")
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    rmode = ssb.SSB_RENDER_RGB if VARIANT == "rgb" else ssb.SSB_RENDER_SPECTRAL
    tiles = args.shard == "tiles" and world > 1
    if tiles:
        # option A of SURVEY 8(e): the SAME frame (spp unchanged) split into interleaved bands of 8 rows (ssb_options.band_*,
        # the reference's tile edge) — strong scaling; every rank gets a share of the expensive and of the cheap rows in ONE
        # launch sequence, and the reduce sums disjoint pixels (+0 elsewhere): bit-identical to one GPU
        total_spp = SPP
        band_opts = [host.options_for(color, W, H, total_spp, seed=1, render_mode=rmode, keep_accumulator=1, prebaked_textures=int(args.prebake),
                                      band_height=8, band_count=world, band_index=rank)]
        opt = band_opts[0]
    else:
        total_spp = SPP * world  # weak scaling: the job is the same frame at spp 64*N
        opt = host.options_for(color, W, H, total_spp, seed=1, sample_begin=rank * SPP, sample_end=(rank + 1) * SPP, render_mode=rmode,
                               prebaked_textures=int(args.prebake))
    npix = W * H

    def device_accum_tensor():
        ptr, count = ctx.accum_device()

        class _Arr:
            __cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Arr(), device=f"cuda:{local_rank}")

    def render_part():
        if tiles:
            for bo in band_opts:            # this rank's row bands, accumulated into one (cleared) buffer
                ctx.render(bo)
        else:
            ctx.render(opt)                 # trace + in-order accumulate, async on the torch stream

    def step_device():
        ctx.clear()
        render_part()
        if world > 1:
            dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)  # the single exchange step: f64 XYZA accumulators
        if rank == 0:
            ctx.resolve_device(opt)         # XYZA (f64) + sRGBA (f32) framebuffers, left on the device

    # warm-up (also allocates the accumulator so that it can be wrapped as a tensor)
    ctx.clear(); ctx.render(opt); ctx.synchronize()
    accum_t = device_accum_tensor() if world > 1 else None
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trace_ms, launches = 0.0, 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    # per-launch duration of the dominant kernel, CUDA events on the launching stream (one more, separately timed step)
    trace_list = []
    for _ in range(3):
        ctx.clear(); ctx.render(opt)
        st = ctx.stats()
        trace_list.append(st.trace_ms)
        launches_per_render = st.launches
    trace_ms = sum(trace_list) / len(trace_list)
    launches = (launches_per_render * (len(band_opts) if tiles else 1) + (1 if rank == 0 else 0)) * args.steps

    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
    dev_ms, wall_ms = t.tolist()
    samples_per_step = npix * SPP * (1 if tiles else world)
    value = samples_per_step * args.steps / (dev_ms * 1e-3) / 1e6

    # ---------------- e2e: host buffers in, host buffers out, every step
    flat_scene = scene.flat
    e2e_scene = ssb.ssb_scene()
    C.memmove(C.byref(e2e_scene), C.byref(flat_scene), C.sizeof(e2e_scene))
    tex_bytes = 0
    if flat_scene.ntextures:  # the textured scenes re-upload their 4096^2 texture from pinned memory every step
        tex_np = host.load_png_rgb8(os.path.join(color.data_root, "data", "scenes", "crystal-lizard-4096.png"))
        pinned_tex = torch.empty(tex_np.shape, dtype=torch.uint8, pin_memory=True)
        pinned_tex.numpy()[...] = tex_np
        tex_desc = (ssb.ssb_texture * 1)()
        tex_desc[0].rgb8 = C.cast(pinned_tex.data_ptr(), C.POINTER(C.c_uint8))
        tex_desc[0].width, tex_desc[0].height = tex_np.shape[1], tex_np.shape[0]
        e2e_scene.textures = tex_desc
        tex_bytes = tex_np.nbytes
    xyza_host = torch.empty((H, W, 4), dtype=torch.float64, pin_memory=True)
    srgba_host = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True)
    xyza_np, srgba_np = xyza_host.numpy(), srgba_host.numpy()
    accum_host = torch.empty(npix * 4, dtype=torch.float64, pin_memory=True) if world > 1 else None
    h2d = tex_bytes + 16 * 1024  # texture + (scene blob + colour tables, < 16 KiB)
    d2h = xyza_np.nbytes + srgba_np.nbytes

    def step_e2e():
        ctx.upload_color(color.flat)
        ctx.upload_scene_async(e2e_scene)  # pinned texels: the copy overlaps the camera-ray stage of the render below
        if world == 1:
            ctx.render_frame(opt, xyza=xyza_np, srgba=srgba_np)
        else:
            ctx.clear(); render_part()
            dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                ctx.resolve(opt, xyza=xyza_np, srgba=srgba_np)
            else:
                ctx.synchronize()

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e2e_steps = max(3, args.steps // 2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    tw = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_wall = (time.perf_counter() - tw) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), e2e_wall)  # host-side work (blob packing) is part of the call: use the larger
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item()
    e2e_value = samples_per_step * e2e_steps / (e2e_ms * 1e-3) / 1e6

    # ---------------- frame checksum of the job just measured (every line): sha256 of the f64 XYZA accumulators on rank 0
    import hashlib

    def accum_sha():
        return hashlib.sha256(np.ascontiguousarray(ctx.read_accum(W, H)).tobytes()).hexdigest()[:16]

    step_device()
    torch.cuda.synchronize()
    frame_sha = accum_sha() if rank == 0 else None

    # ---------------- strong scaling of the FIXED headline job (N > 1): the same spp-SPP frame split over the ranks, both ways
    strong = None
    if world > 1 and not tiles:
        strong = {}
        for mode in ("tiles", "samples"):
            if mode == "tiles":
                so = host.options_for(color, W, H, SPP, seed=1, render_mode=rmode, keep_accumulator=1, prebaked_textures=int(args.prebake),
                                      band_height=8, band_count=world, band_index=rank)
            else:
                so = host.options_for(color, W, H, SPP, seed=1, render_mode=rmode, keep_accumulator=1, prebaked_textures=int(args.prebake),
                                      sample_begin=SPP * rank // world, sample_end=SPP * (rank + 1) // world)
            empty = so.sample_end != 0 and so.sample_end <= so.sample_begin  # more ranks than samples
            full = host.options_for(color, W, H, SPP, seed=1, render_mode=rmode)

            def step_strong(reduce=True):
                ctx.clear()
                if not empty:
                    ctx.render(so)
                if reduce:
                    dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)
                    if rank == 0:
                        ctx.resolve_device(full)

            for _ in range(3):
                step_strong()
            torch.cuda.synchronize(); dist.barrier()
            times = {}
            for key, red in (("frame", True), ("render_only", False)):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(args.steps):
                    step_strong(red)
                b.record(stream)
                torch.cuda.synchronize()
                tt = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=f"cuda:{local_rank}")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dist.barrier()
                times[key] = tt.item() / args.steps
            step_strong()
            torch.cuda.synchronize()
            sha = accum_sha() if rank == 0 else None
            strong[mode] = {"value": npix * SPP / (times["frame"] * 1e-3) / 1e6, "unit": METRIC, "ms_per_frame": times["frame"],
                            "render_only_ms": times["render_only"], "reduce_and_resolve_ms": times["frame"] - times["render_only"],
                            "frame_sha": sha, "scaling": "strong",
                            "bit_identical_to_1gpu": (sha == HEADLINE_SHA) if (mode == "tiles" and (SCENE, VARIANT, W, H, SPP) == ("cornell-srgb", "ours1931", 512, 512, 64) and rank == 0) else None,
                            "shard": "interleaved bands of 8 rows (ssb_options.band_*)" if mode == "tiles" else "sample ranges"}
        dist.barrier()

    if rank == 0:
        peak, peak_src = measured_peaks()
        algo_per_sample = ALGO_BYTES_PER_SAMPLE[SCENE] + (JH_EXTRA_BYTES_PER_SAMPLE[SCENE] if VARIANT == "jh" else 0.0)
        algo_bytes = algo_per_sample * npix * SPP  # per trace-kernel launch (one rank's launch)
        achieved = algo_bytes / (trace_ms * 1e-3) / 1e9
        traffic, issue = None, None
        tp = os.path.join(ROOT, "profiles", "trace_kernel_traffic.json")
        headline = (SCENE, VARIANT, W, H, SPP) == ("cornell-srgb", "ours1931", 512, 512, 64)  # what the ncu capture ran
        traffic_note = None
        if os.path.exists(tp) and headline:
            try:
                prof = json.load(open(tp))
                if prof.get("source_sha") != source_sha():
                    traffic_note = (f"profiles/trace_kernel_traffic.json was captured on another build (source_sha {prof.get('source_sha')} != "
                                    f"{source_sha()}): traffic / issue not quoted")
                    raise ValueError(traffic_note)
                traffic = prof.get("dram_bytes_per_launch")
                wi, ti = prof.get("warp_instructions_per_frame"), prof.get("thread_instructions_per_frame")
                if wi and clocks.get("sm_mhz"):
                    # the bound that actually binds (SURVEY 8(d) "FP32-issue bound"): warp instructions of one frame (ncu
                    # smsp__inst_executed.sum over its launches, profiles/) against the SMs' issue slots during the measured step
                    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
                    slots = sms * 4 * clocks["sm_mhz"] * 1e6 * (dev_ms / args.steps) * 1e-3
                    issue = {"warp_instructions_per_frame": wi, "thread_instructions_per_sample": ti / (npix * SPP),
                             "lanes_per_instruction": ti / wi, "issue_slots_per_frame": slots, "frac": wi / slots,
                             "note": "fraction of all warp-issue slots (SMs x 4 schedulers x SM clock x step time) used by the frame's instructions"}
            except Exception:
                traffic, issue = None, None
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if tiles else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "reference scene (hard-coded geometry, shipped spectra + 4096^2 sRGB texture); per-sample seeded RNG",
            "config": {"workload": f"{SCENE} {W}x{H} spp{SPP} per GPU (job spp {total_spp}), hero-wavelength x4, variant {VARIANT} (upsampling + observer), "
                                   f"ELS on, MAX_DEPTH 10" + (", prebaked JH coefficient textures" if args.prebake else ""), "parallelism": (f"row-band tiles x{world}, one NCCL reduce of f64 XYZA" if tiles else f"sample-sharded x{world}, one NCCL reduce of f64 XYZA") if world > 1 else "single GPU",
                       "l2": "no flush needed: every step streams ~8 GB of path records / fold records through HBM (>> 126 MB L2)",
                       "timing": "CUDA events on the launching stream around K steps, max over ranks", "wall_ms_per_step": wall_ms / args.steps},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "calls": "ssb_upload_color + ssb_upload_scene_async (pinned RGB8 texture) + ssb_render_frame -> pinned XYZA f64 + sRGBA f32"},
            "gpu_launches": int(launches),
            "frame_sha": frame_sha,
            "frame_sha_note": "sha256[:16] of the f64 XYZA accumulators of the measured job on rank 0" +
                              (f"; the 1-GPU spp{SPP} frame of the headline config is {HEADLINE_SHA}" if (SCENE, VARIANT, W, H, SPP) == ("cornell-srgb", "ours1931", 512, 512, 64) else ""),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src,
                         "kernel": "bounce stage = ssb_intersect_kernel + counting sort + ssb_shade_kernel over all path depths of one frame "
                                   "(CUDA events around the launch sequence)", "kernel_ms": trace_ms,
                         "algorithmic_bytes_per_sample": algo_per_sample, "issue": issue, "traffic_note": traffic_note,
                         "note": "achieved = SURVEY.md 8(d) algorithmic bytes (wavefront ray-state model) / bounce-stage time; traffic = ncu dram bytes of the "
                                 "same launches (profiles/). The stage is instruction-issue bound (un-fused fp32 + f64 exact libm), DRAM ~20-30 % busy: see DESIGN.md (d)"},
        }
        if strong is not None:
            # the fixed spp-SPP frame at N GPUs; the tiles frame must be the 1-GPU frame bit for bit (per-sample seeding, disjoint pixels)
            line["strong"] = strong
            line["strong"]["job"] = f"{SCENE} {W}x{H} spp{SPP} (fixed), split over {world} GPUs + one NCCL reduce of f64 XYZA + resolve on rank 0"
            line["strong"]["limiter"] = ("per-GPU slice of %.2f ms against a floor of ~30 dependent kernel launches per pass (each depth's queue length is "
                                         "read on the device, so the launches cannot be merged) + the 8 MB ncclReduce and resolve: see reduce_and_resolve_ms"
                                         % strong["tiles"]["render_only_ms"])
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(W, H, args.cpu_spp)
            except SystemExit as e:
                line["cpu_baseline"] = {"value": None, "unit": METRIC, "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}
        print(json.dumps(line), file=JSON_OUT or sys.stdout, flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


def main():
    global SCENE, VARIANT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)  # 100 frames x ~17 ms: a timed region of > 1.5 s by default
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--cpu-spp", type=int, default=16, help="spp of the bounded CPU-baseline sample")
    ap.add_argument("--ref-spp", type=int, default=16, help="spp per step of the reference arm (same bounded sample as cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", default="samples", choices=["samples", "tiles"],
                    help="multi-GPU decomposition: sample ranges (default, weak scaling) or row bands of the same frame (strong scaling)")
    ap.add_argument("--scene", default=SCENE, choices=sorted(ALGO_BYTES_PER_SAMPLE),
                    help="default = BASELINE configs[1]; the others are SURVEY 8(d) C3-C5 (not the headline)")
    ap.add_argument("--prebake", action="store_true",
                    help="variant jh only: ssb_options.prebaked_textures (Jakob-Hanika coefficient textures, baked once per upload)")
    ap.add_argument("--variant", default=VARIANT, choices=["ours1931", "ours2006", "meng", "jh", "rgb"])
    args = ap.parse_args()
    SCENE, VARIANT = args.scene, args.variant
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return main_reference(args, rank, world)
    return main_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
