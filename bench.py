#!/usr/bin/env python3
"""bench.py -- the benchmarks of the path-tracing hot path on B200 (BASELINE.json configs C1..C5).

  python bench.py [--gpus N --steps K --warmup W]                    headline: config C2, one process per GPU under torchrun
  python bench.py --config c1|c2|c3|c4|c5 [--scaling weak|strong]    the other BASELINE configs, same JSON contract
  python bench.py --impl reference [--config ...]                    the reference's own kernel arithmetic on the host CPU cores

One STEP = one full render of the configured workload.  `--scaling weak` (default): every rank renders the configured spp on
its own, disjoint, sample-index range.  `--scaling strong` (default for c5): the configured spp is the TOTAL, split over the
ranks by sample index (yune_shard_samples), one SUM-reduce of the fp32 accumulation buffers to rank 0, rank 0 tonemaps.

Prints ONE JSON line (rank 0).  `value` = Msamples/s with inputs resident in HBM, timed on the device with CUDA events (max
over ranks); `e2e` = the same metric through the C ABI with pinned HOST buffers (scene upload + render + tonemap + image
read-back, wall clock); `roofline` = the trace kernel's steady-state launches against the measured HBM bandwidth with the
oracle's work model; `roofline_onchip` = the same launches against the SM issue rate (the bound that is physical
for the cache-resident scenes); `roofline_shade` = the shade kernel; `cpu_baseline` = the reference's kernels (oracle/_ref,
else the restatement) on this box's cores, bounded sample.  The oracle is only ever used here as the CPU baseline / work
counter, never on the measured GPU path.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 12345

# BASELINE.json configs.  `ref_variant`: which compiled reference kernel can time this config on the CPU
# (None: only the restatement can -- the reference has a single built-in light, no Oren-Nayar, and a 1500-entry traversal queue).
CONFIGS = {
    "c1": dict(scene="cornellbox", kernel="udpt.cl", opts="", W=512, H=512, spp=64, variant="udpt", ref_variant="udpt",
               metric="Msamples/s (Cornell box 512x512, udpt NEE+RR, 64 spp)",
               workload="C1: cornellbox (60 tris, 15 BVH nodes), udpt.cl (NEE + Russian roulette, MIS off), 512x512x64spp",
               cpu_sample=(256, 256, 32)),
    "c2": dict(scene="teapot", kernel="udpt.cl", opts="-DMIS", W=1024, H=1024, spp=1024, variant="udpt_mis", ref_variant="udpt_mis", transmissive=True,
               metric="Msamples/s (Cornell-teapot 1024x1024, transmissive teapot, udpt+MIS, 1024 spp)",
               workload="C2: cornellbox-teapot (6380 tris, 2199 BVH nodes), teapot=mirror-transmissive, udpt.cl -DMIS, 1024x1024x1024spp",
               cpu_sample=(256, 256, 32)),
    "c3": dict(scene="teapot", kernel="bdpt.cl", opts="", W=1024, H=1024, spp=512, variant="bdpt", ref_variant=None, two_lights=True, oren_nayar=0.25,
               metric="Msamples/s (Cornell-teapot 1024x1024, naive BDPT, 2 lights, Oren-Nayar walls + Phong teapot, 512 spp)",
               workload="C3: cornellbox-teapot, bdpt.cl (BDPT_BOUNCES 20), 2 quad lights, Oren-Nayar walls (sigma^2 0.25) + modified-Phong teapot, 1024x1024x512spp",
               cpu_sample=(128, 128, 8)),
    "c4": dict(scene="c4", kernel="udpt.cl", opts="", W=3840, H=2160, spp=256, variant="udpt", ref_variant=None,
               metric="Msamples/s (synthetic 10.5 M triangles, 3840x2160, udpt, 256 spp)",
               workload="C4: Cornell walls + 2 icospheres at subdivision 9 (10,485,800 tris, 3.1 M reference BVH nodes), udpt.cl, 3840x2160x256spp",
               cpu_sample=(192, 108, 4)),
    "c5": dict(scene="teapot", kernel="udpt.cl", opts="-DMIS", W=3840, H=2160, spp=16384, variant="udpt_mis", ref_variant="udpt_mis", transmissive=True, strong=True,
               metric="Msamples/s (Cornell-teapot 3840x2160, udpt+MIS, 16384 spp sample-partitioned over the GPUs)",
               workload="C5: C2 scene, udpt.cl -DMIS, 3840x2160, 16384 spp in total split by sample index, one SUM reduce of the 132.7 MB accumulation buffer",
               cpu_sample=(256, 144, 16)),
}


def build_scene(cfg):
    """(tris, mats, nodes, lights or None) of a config, reference record layouts."""
    from tests.refbind import load_golden_scene
    import yune_b200 as yb
    if cfg["scene"] == "c4":
        from yune_b200.scenes import synthetic_c4
        tris, mats, _ = load_golden_scene("cornellbox")
        sc = yb.Scene().setGeometry(synthetic_c4(tris, 9), mats)       # the product's host builder: byte-identical to the reference's
        return np.array(sc.vert_data), np.array(sc.mat_data), np.array(sc.bvh), None
    tris, mats, nodes = load_golden_scene(cfg["scene"])
    tris, mats = tris.copy(), mats.copy()
    if cfg.get("transmissive"):
        tris["matID"] = np.where(tris["matID"] == 3, 4, tris["matID"])   # 'usemtl teapot' -> 'mirror-transmissive' (config C2)
    lights = None
    if cfg.get("two_lights"):
        lights = np.concatenate([yb.LIGHT_BDPT, yb.quad_light((0.6, 0.0, -3.6), (-1, 0, 0), (8, 8, 8), (0, 0.3, 0), (0, 0, 0.3))])
    if cfg.get("oren_nayar"):
        mats["alpha_x"] = cfg["oren_nayar"]
    return tris, mats, nodes, lights


def load_scene():
    """The C2 scene (kept for the tools that import it)."""
    return build_scene(CONFIGS["c2"])[:3]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for i, n in enumerate(names):
                    if r[5 + i].lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_run(cfg, scene, width, height, spp, threads):
    """The reference's kernel arithmetic on host cores for one bounded sample of the config: oracle/_ref (its own .cl text compiled
    as C++) when that kernel can express the config, else the restatement.  Returns (Msamples/s, kind)."""
    from tests.refbind import RefKernels, Oracle, default_cam_array, frame_rands, have_ref
    tris, mats, nodes, lights = scene
    cam = default_cam_array()
    rands = frame_rands(SEED, spp)
    t0 = time.perf_counter()
    if have_ref() and cfg["ref_variant"]:
        RefKernels().render(cfg["ref_variant"], cam, tris, mats, nodes, width, height, rands, threads=threads)
        kind = "reference"
    else:
        oc = Oracle.config(cfg["variant"], threads=threads, lights=lights, oren_nayar=1 if cfg.get("oren_nayar") else 0,
                           heap_size=0 if cfg["scene"] == "c4" else -1)
        Oracle().render(oc, cam, tris, mats, nodes, width, height, rands, lights=lights)
        kind = "port"
    dt = time.perf_counter() - t0
    return width * height * spp / dt / 1e6, kind


def reference_opencl_run(cfg, scene, width, height, frames, want_image=False):
    """SURVEY.md 8c / 8d "opportunistic" second baseline: the reference's own udpt.cl run ON THE GPU by NVIDIA's OpenCL driver, with
    the reference's launch protocol (oracle/_ref/yune_ref_ocl, built from the reference tree by oracle/gen_ref_ocl.py; test /
    measurement infrastructure).  One frame = one sample per pixel.  Returns the binary's JSON line as a dict (or
    {"unavailable": why}); with want_image also the final running-mean image."""
    import tempfile
    from tests.refbind import default_cam_array, frame_rands
    exe = os.path.join(ROOT, "oracle", "_ref", "yune_ref_ocl")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/yune_ref_ocl not built (needs /root/reference at build time)"}
    if cfg["ref_variant"] is None:
        return {"unavailable": "the reference kernel cannot express this config (one built-in light, no Oren-Nayar, 1500-entry queue)"}
    tris, mats, nodes, lights = scene
    with tempfile.TemporaryDirectory() as d:
        for name, a in (("tris", tris), ("mats", mats), ("nodes", nodes), ("cam", default_cam_array())):
            np.ascontiguousarray(a).tofile(os.path.join(d, name + ".bin"))
        np.array(frame_rands(SEED, frames + 1), np.uint32).tofile(os.path.join(d, "rands.bin"))
        try:
            p = subprocess.run([exe, d, str(width), str(height), str(frames), "1", cfg["opts"]], capture_output=True, text=True, timeout=300)
            out = json.loads(p.stdout.strip().splitlines()[-1])
        except Exception as e:
            return {"unavailable": "yune_ref_ocl failed: %s" % e}
        if want_image and "unavailable" not in out:
            out["image"] = np.fromfile(os.path.join(d, "image.bin"), np.float32).reshape(height, width, 4)
    return out


def cpu_sample_text(cfg, w, h, spp, kind):
    why = ""
    if kind == "port" and cfg["ref_variant"] is None:
        why = " (restatement: the reference kernel has one built-in light, no Oren-Nayar and a 1500-entry traversal queue)"
    return "%dx%dx%dspp of the %s workload, reference RNG%s" % (w, h, spp, cfg["name"].upper(), why)


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests.refbind import ncores
    scene = build_scene(cfg)
    cores = ncores()
    w, h, spp = cfg["cpu_sample"]
    spp = max(spp // 2, 1)
    for _ in range(args.warmup):
        cpu_reference_run(cfg, scene, 64, 36 if cfg["W"] != cfg["H"] else 64, 2, cores)
    t0 = time.perf_counter()
    kind = "port"
    for _ in range(args.steps):
        _, kind = cpu_reference_run(cfg, scene, w, h, spp, cores)
    dt = time.perf_counter() - t0
    value = w * h * spp * args.steps / dt / 1e6
    sample = cpu_sample_text(cfg, w, h, spp, kind) + " per step"
    print(json.dumps({
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (reference scene buffers, default camera)",
        "config": {"workload": cfg["workload"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def algorithmic_bytes_per_ray(r, tris, nodes, warm_spp):
    """SURVEY.md 8d work model: B_ray = 32*N_box + 48*N_tri + 48, N_* counted by the ORACLE's ordered stack walk with
    t-pruning over the reference BVH, on rays captured from one steady-state wavefront iteration of this very render."""
    from tests.refbind import Oracle
    o = Oracle()
    r.captureRays(24, 1 << 17)          # iteration 24: pool full, path mix stationary
    r.samples_taken = 0
    r.enqueueKernels(warm_spp, reset=True)
    out = {}
    for which, name in ((0, "extend"), (1, "shadow")):
        od, tm, nq = r.readCaptured(which)
        if od.shape[0] == 0:
            out[name] = {"rays": 0, "box": 0.0, "tri": 0.0, "bytes": 48.0}
            continue
        nb, nt = o.count_work(od, tm, which, tris, nodes)
        n = od.shape[0]
        out[name] = {"rays": n, "box": nb / n, "tri": nt / n, "bytes": 32.0 * nb / n + 48.0 * nt / n + 48.0}
    r.captureRays(0, 0)
    return out


def frame_times(m, lib, ctx, W, H):
    """The reference's interactive call pattern: ONE sample per pixel per yune_render (src/RendererCore.cpp:483-486).  ms per call
    when every call returns a complete image (default) and with option "pipeline" (a call returns when its samples are handed
    out, paths in flight are carried into the next call; yune_finish completes the image)."""
    import yune_b200 as yb
    m.setOption("time_stages", 0)
    st = yb._native.Stats()
    exact = []
    for f in range(6):
        m.check(lib.yune_render(ctx, f, 1, 1, SEED, 1 if f == 0 else 0)); lib.yune_get_stats(ctx, C.byref(st)); exact.append(st.render_ms)
    out = {"spp_per_call": 1, "image": "%dx%d" % (W, H), "complete_image_per_call_ms": float(np.median(exact[1:]))}
    try:
        m.setOption("pipeline", 1)
        n = 64
        m.check(lib.yune_render(ctx, 0, 1, 1, SEED, 1))
        t0 = time.perf_counter()
        for f in range(1, n + 1):
            m.check(lib.yune_render(ctx, f, 1, 1, SEED, 0))
        out["pipelined_ms"] = (time.perf_counter() - t0) * 1e3 / n
        lib.yune_get_stats(ctx, C.byref(st)); out["pipelined_paths_in_flight"] = int(st.carried_paths)
        m.check(lib.yune_finish(ctx)); lib.yune_get_stats(ctx, C.byref(st)); out["finish_ms"] = st.finish_ms
        out["pipelined_pool_slots"] = int(m.getOption("pool_slots_in_use"))
    finally:
        m.setOption("pipeline", 0)
    return out


def profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    try:
        return json.load(open(p))
    except Exception:
        return None


def run_ours(args, cfg):
    import torch
    import yune_b200 as yb
    from yune_b200.dist import shard_samples, reduce_sum_to_root, sum_buffer_as_tensor
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line, the JSON: whatever libraries write to file descriptor 1 meanwhile (NCCL prints its
    # version there) is sent to stderr, and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def pinned(a):
        """A page-locked copy of a numpy array (same dtype / shape), so that the e2e uploads and the read-back run from pinned host memory."""
        t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8).pin_memory()
        v = t.numpy()[:a.nbytes].view(a.dtype).reshape(a.shape)
        v[...] = a
        return v, t

    W, H, SPP = cfg["W"], cfg["H"], cfg["spp"]
    strong = args.scaling == "strong"
    tris, mats, nodes, lights = build_scene(cfg)
    (tris, _k1), (mats, _k2), (nodes, _k3) = pinned(tris), pinned(mats), pinned(nodes)
    cam = yb.default_camera()
    m = yb.CUDAManager().setup(local)
    r = yb.RendererCore(m, W, H)
    r.seed = SEED

    def upload():
        assert m.createRenderProgram(cfg["kernel"], compiler_opts=cfg["opts"]), m.last_message
        assert m.createPostProcProgram("tonemap.cl"), m.last_message
        assert m.setupVertexBuffer(tris) and m.setupMatBuffer(mats) and m.setupBVHBuffer(nodes) and m.setupImageBuffers(W, H), m.last_message
        if lights is not None:
            assert m.setLightSources(lights), m.last_message
        m.setOption("oren_nayar", 1 if cfg.get("oren_nayar") else 0)
        m.setupCameraBuffer(cam)
    upload()
    if strong:
        spp_begin, spp_rank = shard_samples(SPP, rank, world)          # a fixed total, split by sample index
    else:
        spp_begin, spp_rank = rank * SPP, SPP                           # every rank renders the full workload on its own range
    warm_spp = min(spp_rank, 64) if SPP > 2048 else spp_rank           # a 16 k-spp step is minutes long: warm up on the same scene at 64 spp
    if world > 1:                  # the fixed-point buffer exists after the first render of this size
        m.check(r._lib.yune_render(r._ctx, spp_begin, 1, 1, SEED, 1))
    sum_t = sum_buffer_as_tensor(r) if world > 1 else None          # int64 fixed point (option "deterministic", default): exact, order-free reduce
    lib, ctx = r._lib, r._ctx
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step(spp_count):
        m.check(lib.yune_render(ctx, spp_begin, spp_count, 1, SEED, 1))
        st = yb._native.Stats(); lib.yune_get_stats(ctx, C.byref(st))
        extra_ms, reduce_ms = 0.0, 0.0
        if world > 1:                      # the path's single exchange step: SUM-reduce the fp32 accumulation buffers over NVLink
            ev0.record(); reduce_sum_to_root(sum_t); torch.cuda.synchronize()      # the context's stream does not order itself against torch's
            if rank == 0:
                m.check(lib.yune_sum_refresh(ctx)); m.check(lib.yune_synchronize(ctx))      # float view of the reduced integers
            ev1.record(); torch.cuda.synchronize()
            reduce_ms = ev0.elapsed_time(ev1); extra_ms += reduce_ms
        if rank == 0:
            m.check(lib.yune_tonemap(ctx))
            st2 = yb._native.Stats(); lib.yune_get_stats(ctx, C.byref(st2)); extra_ms += st2.tonemap_ms
        return st, extra_ms, reduce_ms

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(warm_spp)
    m.setOption("time_stages", 16)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms, reduce_total = 0.0, 0.0
    keys = ("extend_rays", "shadow_rays", "kernel_launches", "trace_ms", "shade_ms", "timed_iterations", "trace_launches", "iterations",
            "diffuse_visits", "specular_visits", "regenerations", "slot_visits", "steady_iterations", "steady_timed_iterations",
            "steady_extend_rays", "steady_shadow_rays", "steady_trace_ms", "steady_shade_ms")
    agg = {k: 0 for k in keys}
    launches = 0
    for _ in range(args.steps):
        st, extra, red = step(spp_rank)
        dev_ms += st.render_ms + extra; reduce_total += red
        for k in keys:
            agg[k] += getattr(st, k)
        launches += st.kernel_launches + 1 + (1 if rank == 0 else 0)
    barrier()
    pool_in_use = int(m.getOption("pool_slots_in_use"))
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    m.setOption("time_stages", 0)
    my_ms = dev_ms
    if world > 1:
        tmax = torch.tensor([dev_ms], device="cuda"); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dev_ms = float(tmax.item())
    samples_per_step = W * H * (SPP if strong else SPP * world)
    value = samples_per_step * args.steps / (dev_ms * 1e-3) / 1e6

    # ---- end-to-end through the C ABI with host buffers: upload + render + tonemap + read-back, wall clock ----
    ldr, _k4 = pinned(np.zeros((H, W, 4), np.float32))
    e2e_steps = args.steps if SPP <= 2048 else 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        upload()
        step(spp_rank)
        if rank == 0:
            m.check(lib.yune_read_ldr(ctx, ldr.ctypes.data_as(C.c_void_p)))
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_s = float(tt.item())
    e2e_value = samples_per_step * e2e_steps / e2e_s / 1e6
    h2d = int(tris.nbytes + mats.nbytes + nodes.nbytes + 80 + (lights.nbytes if lights is not None else 0))
    d2h = int(ldr.nbytes)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_trace), STEADY-STATE launches only (pool full): rays, bytes, duration and the ncu
    # traffic figure then all describe the same kind of launch ----
    peak, peak_src = measured_peak()
    bdpt = cfg["kernel"] == "bdpt.cl"
    name = cfg["name"]
    prof = profile_json("trace_traffic_%s.json" % name) or (profile_json("trace_traffic.json") if name == "c2" else None) or {}
    sprof = profile_json("shade_traffic_%s.json" % name) or (profile_json("shade_traffic.json") if name == "c2" else None) or {}
    work = algorithmic_bytes_per_ray(r, tris, nodes, min(spp_rank, 64))
    have_steady = agg["steady_timed_iterations"] > 0 and agg["steady_iterations"] > 0
    if have_steady:
        avg_launch_ms = agg["steady_trace_ms"] / agg["steady_timed_iterations"]
        ext_per_launch = agg["steady_extend_rays"] / agg["steady_iterations"]
        sh_per_launch = agg["steady_shadow_rays"] / agg["steady_iterations"]
        shade_ms = agg["steady_shade_ms"] / agg["steady_timed_iterations"]
        which = "steady-state launches (pool full), every 16th timed with CUDA events on the launching stream"
    else:                                   # a job too short to have a steady state (C1): all launches
        avg_launch_ms = agg["trace_ms"] / max(agg["timed_iterations"], 1)
        ext_per_launch = agg["extend_rays"] / max(agg["trace_launches"], 1)
        sh_per_launch = agg["shadow_rays"] / max(agg["trace_launches"], 1)
        shade_ms = agg["shade_ms"] / max(agg["timed_iterations"], 1)
        which = "all launches of the job (no steady state: the job is a few iterations long)"
    bytes_per_launch = ext_per_launch * work["extend"]["bytes"] + sh_per_launch * work["shadow"]["bytes"]
    achieved = bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
    resident = cfg["scene"] != "c4"
    traffic, traffic_how = None, None
    if prof.get("dram_bytes_per_launch") and prof.get("pool_slots", pool_in_use) == pool_in_use:      # same pool size = same kind of launch
        traffic, traffic_how = prof["dram_bytes_per_launch"], "DRAM bytes of the captured steady-state launch (same pool size)"
    elif prof.get("dram_bytes_per_ray"):      # captured at another pool size: the per-ray figure (node / triangle / queue sectors of one ray) times this launch's rays
        traffic = prof["dram_bytes_per_ray"] * (ext_per_launch + sh_per_launch)
        traffic_how = "captured launch's DRAM bytes per ray (%.0f B, pool of %d slots) x the live rays per launch" % (prof["dram_bytes_per_ray"], prof.get("pool_slots", 0))
    roofline = {"bound": "hbm", "kernel": "k_trace", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "hbm_frac_measured": (traffic / (avg_launch_ms * 1e-3) / 1e9 / peak) if traffic and avg_launch_ms > 0 else None, "peak_source": peak_src, "avg_launch_ms": avg_launch_ms, "bytes_per_launch": bytes_per_launch,
                "launches": which, "rays_per_launch": {"extend": ext_per_launch, "shadow": sh_per_launch}, "work_model": work,
                "traffic_source": prof.get("source"), "traffic_how": traffic_how,
                "trace_share_of_step": agg["trace_ms"] / max(agg["trace_ms"] + agg["shade_ms"], 1e-9),
                "note": ("the scene (<= 0.9 MB) is resident in shared memory / L1 / L2: achieved is the EFFECTIVE bandwidth of the work model "
                         "(oracle's ordered walk of the reference BVH), it exceeds the HBM peak by construction and is not a physical fraction; "
                         "the physical bound of this kernel on this config is the issue rate, see roofline_onchip") if resident else
                        "10.5 M triangles (1.2 GB of traversal data): the walk is bound by random-sector DRAM/L2 latency; traffic = measured DRAM bytes of a steady-state launch"}
    # on-chip bound: warp instructions issued per second against 148 SMs x 4 schedulers x clock.  Instructions per ray, lanes per
    # instruction and shared-memory wavefronts per ray come from the committed ncu capture of a steady-state launch of the same
    # build and workload (profiles/); rays per launch and the launch duration are live.
    onchip = None
    if prof.get("warp_inst_per_ray"):
        sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
        issue_peak = 148 * 4 * sm_clock * 1e6
        rays = ext_per_launch + sh_per_launch
        inst_s = rays * prof["warp_inst_per_ray"] / (avg_launch_ms * 1e-3) if avg_launch_ms > 0 else 0.0
        onchip = {"bound": "issue", "kernel": "k_trace", "achieved": inst_s / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-inst/s", "frac": inst_s / issue_peak,
                  "lanes_per_inst": prof.get("lanes_per_inst"), "thread_frac": inst_s / issue_peak * (prof.get("lanes_per_inst") or 32.0) / 32.0,
                  "warp_inst_per_ray": prof["warp_inst_per_ray"], "smem_wavefronts_per_s": (rays * prof.get("smem_wavefronts_per_ray", 0.0) / (avg_launch_ms * 1e-3)) if avg_launch_ms > 0 else None,
                  "smem_wavefront_peak_per_s": 148 * sm_clock * 1e6, "ncu_issue_active_pct": prof.get("issue_active_pct"),
                  "source": prof.get("source"), "note": "achieved = live rays per steady-state launch x ncu warp instructions per ray / live launch duration"}

    # ---- second kernel of the step: the shade kernel streams the path pool: algorithmic bytes per slot visit = the state a visit
    # of that kind must read and write (16-byte fields), counted by the kernel itself per kind ----
    roofline_shade = None
    if not bdpt:
        B_CLASSIFY, B_DIFFUSE, B_SPECULAR, B_REGEN = 20.0, 330.0, 260.0, 100.0
        n_l = max(agg["trace_launches"], 1)
        shade_bytes = (agg["slot_visits"] * B_CLASSIFY + agg["diffuse_visits"] * B_DIFFUSE + agg["specular_visits"] * B_SPECULAR + agg["regenerations"] * B_REGEN) / n_l
        shade_avg_ms = agg["shade_ms"] / max(agg["timed_iterations"], 1)
        shade_ach = shade_bytes / (shade_avg_ms * 1e-3) / 1e9 if shade_avg_ms > 0 else 0.0
        roofline_shade = {"bound": "hbm", "kernel": "k_shade_dense", "achieved": shade_ach, "peak": peak, "unit": "GB/s", "frac": shade_ach / peak,
                          "traffic": sprof.get("dram_bytes_per_launch"), "traffic_source": sprof.get("source"), "avg_launch_ms": shade_avg_ms, "steady_launch_ms": shade_ms,
                          "bytes_per_launch": shade_bytes, "launches": "job average (visit counters are per job); traffic is a steady-state launch",
                          "visits_per_launch": {"slots": agg["slot_visits"] / n_l, "diffuse": agg["diffuse_visits"] / n_l, "specular": agg["specular_visits"] / n_l, "regenerate": agg["regenerations"] / n_l},
                          "bytes_per_visit": {"classify": B_CLASSIFY, "diffuse": B_DIFFUSE, "specular": B_SPECULAR, "regenerate": B_REGEN}}

    # ---- CPU baseline: the reference's kernel on this box's cores, bounded sample of the same workload (N = 1 only) ----
    from tests.refbind import ncores
    cores = ncores()
    cw, ch, cspp = cfg["cpu_sample"]
    cpu_v, kind = cpu_reference_run(cfg, (tris, mats, nodes, lights), cw, ch, cspp, cores) if world == 1 else (None, "reference")
    rays = agg["extend_rays"] + agg["shadow_rays"]
    # the reference's kernel on this very GPU (NVIDIA OpenCL), bounded: 16 frames of the config's image size
    ref_ocl, frames = None, None
    if world == 1:                       # side measurements: never allowed to cost the bench line
        try:
            ref_ocl = reference_opencl_run(cfg, (tris, mats, nodes, lights), W, H, 16 if W * H <= (1 << 21) else 4)
            ref_ocl["sample"] = "%dx%d, %d frames of 1 spp, reference RNG" % (W, H, ref_ocl.get("frames", 0)) if "unavailable" not in ref_ocl else None
        except Exception as e:
            ref_ocl = {"unavailable": "reference_opencl_run raised %r" % (e,)}
        try:
            frames = frame_times(m, lib, ctx, W, H) if not bdpt and W * H <= (1 << 21) else None
        except Exception as e:
            frames = {"unavailable": "frame_times raised %r" % (e,)}
    line = {
        "metric": cfg["metric"], "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (reference scene buffers, default camera, counter-based RNG seed 12345)",
        "config": {"workload": cfg["workload"], "name": name, "spp_per_rank": spp_rank, "sample_range_of_rank_0": [spp_begin, spp_begin + spp_rank],
                   "split": "total spp split over the ranks by sample index (yune_shard_samples)" if strong else "every rank renders the configured spp on its own sample-index range",
                   "pool_slots": pool_in_use, "warmup_spp": warm_spp,
                   "l2": "per-step working set (path pool + queues ~350 B/slot, %.1f GB at %d slots; %.1f MB accumulation) exceeds the 126 MB L2" % (350.0 * pool_in_use / 1e9, pool_in_use, W * H * 16 / 1e6)},
        "mrays_per_s": rays / (my_ms * 1e-3) / 1e6 * world,
        "rays_per_sample": rays / (W * H * spp_rank * args.steps),
        "wall_s_timed_region": wall_s,
        "reduce_ms_per_step": reduce_total / args.steps if world > 1 else None,
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_onchip": onchip,
        "roofline_shade": roofline_shade,
        "reference_opencl_same_gpu": ref_ocl,
        "frame_ms": frames,
        "cpu_baseline": None if world > 1 else {"value": cpu_v, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": cpu_sample_text(cfg, cw, ch, cspp, kind)},
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config], name=args.config)
    if args.scaling is None:
        args.scaling = "strong" if cfg.get("strong") else "weak"
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
