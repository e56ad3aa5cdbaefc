#!/usr/bin/env python
"""bench.py -- Mrays/s of the Mallie render hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[3], the one the north-star target is quoted on):
  procedurally tessellated bumpy sphere, N=500 -> exactly 1 000 000 triangles, 1920x1080, 16 spp,
  primary closest-hit ray + one shadow (occlusion) ray per primary hit, camera eye (0,0,3) -> origin.
A "step" is one 16-spp frame: 33.2 M primary rays + ~11.4 M shadow rays.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* value : whole-job Mrays/s with everything resident in HBM (framebuffer stays on the device).
          N > 1: image rows are interleaved across ranks in bands of 4 scanlines (strong scaling, the
          frame is fixed), each rank renders its bands, one NCCL all-gather of the framebuffer per frame.
* e2e   : the same frame through the C-ABI call a Mallie host makes (mb200_render_frame) with pinned HOST
          image / count buffers; the device->host copy of the framebuffer is inside the timed region.
* roofline : the traversal kernel k_trace_sm (closest-hit launches over camera rays + any-hit launches over shadow
          rays; per-instantiation breakdown under roofline.kernels).  achieved = algorithmic bytes per launch
          (64 B/node popped + 88 B/triangle tested + 48 B ray + 32 B hit record, counted by the CPU oracle in
          reference traversal order for the exact ray set; a shadow ray counts as the closest-hit Traverse that
          defines its oracle) / average launch duration, timed with CUDA events around every launch on its
          stream inside the timed region (mb200_scene_timing), against the measured HBM copy bandwidth in
          MEASURED_PEAKS.json.  Consecutive batches of a frame run on two streams so that one launch's drain
          phase is filled by the next launch: durations are the UNION of the launches' time spans (per
          instantiation, and over both for the headline figure) divided by the number of launches.
          traffic = ncu dram read+write bytes per launch (profiles/r1_traffic.json).
* cpu_baseline : the unmodified reference (oracle/_ref) tracing a sample of the same ray set on the host cores.
* --impl reference : times the reference's own OpenMP CPU path on the same workload (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, SPP, SPHERE_N = 1920, 1080, 16, 500
EYE, LOOKAT, LIGHT = (0.0, 0.0, 3.0), (0.0, 0.0, 0.0), (2.0, 4.0, 3.0)
BAND_ROWS = 4          # one tile row per band: the finest interleave (rank load differs by < 1 band in ~22)
METRIC = "Mrays/s primary+shadow at 1920x1080"
L2_FLUSH_BYTES = 512 << 20


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print to stdout (NCCL's version banner, ...) goes to stderr; the one JSON line of the
    contract is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def config_dict(n_gpus):
    return {
        "workload": "bumpy-sphere N=500 (1,000,000 triangles, 501,501 vertices), 1920x1080, 16 spp, "
                    "primary closest-hit + 1 shadow ray per hit, eye (0,0,3) lookat (0,0,0) fov 45, light (2,4,3)",
        "triangles": 1000000, "resolution": [W, H], "spp": SPP, "shader": "primary+shadow",
        "parallelism": "1 GPU" if n_gpus == 1 else f"image rows in {BAND_ROWS}-scanline bands interleaved over "
                                                    f"{n_gpus} GPUs, full scene replica per GPU, NCCL all-gather of the framebuffer",
        "l2": "L2 flushed between timed steps (512 MiB memset); scene (84 MB) is L2-resident within a step",
    }


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The sampler is
    started before the warm-up (nvidia-smi needs ~0.2 s to deliver its first line) and only the samples whose
    timestamp falls inside [mark_start(), mark_end()] are used; a timed region shorter than the sampling period
    falls back to the samples taken under the same load during the warm-up and says so."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        # nvidia-smi needs a moment (longer when several ranks start at once) before its first line: wait for it,
        # so that short timed regions are covered
        t_end = time.time() + 5.0
        while self.p is not None and time.time() < t_end and os.path.getsize(self.f.name) == 0:
            time.sleep(0.02)

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[1]), float(c[2]), float(c[3]),
                             [nm for k, nm in enumerate(names) if c[5 + k].lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        window = "timed region"
        if not inside:          # timed region shorter than one sampling period: samples since the warm-up began
            inside, window = rows, "warm-up + timed region (timed region shorter than the sampling period)"
        if inside:
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)),
                       reasons=sorted({nm for r in inside for nm in r[4]}), samples=len(inside),
                       power_w_max=float(max(r[3] for r in inside)), window=window)
        return out


def build_inputs():
    from mallie_b200.procedural import bumpy_sphere
    return bumpy_sphere(SPHERE_N)


# ----------------------------------------------------------------------------------------------------
# CPU side (oracle / reference): test infrastructure, used here only as checker + baseline
# ----------------------------------------------------------------------------------------------------
class CpuSide:
    def __init__(self, v, f):
        from oracle import orabind as O
        from oracle import refbind as R
        self.O, self.R = O, R
        self.mesh = O.Mesh(v, f)
        self.bvh = O.BVH.build(self.mesh)
        self.frame = O.camera_frame(EYE, LOOKAT, width=W, height=H)
        self.ref = None
        if R.available():
            self.ref = R.RefScene.from_arrays(v, f)
            self.ref.build()
        # all host threads, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def pass_rays(self, k):
        """Exact ray set of pass k (jittered camera rays + shadow rays) and its oracle node/triangle counts."""
        _, _, info = self.bvh.render_pass(self.frame, W, H, rng_mode=1, pass_index=k, shader=1, light=LIGHT,
                                          emit_rays=True, nthreads=self.cores)
        rays = np.concatenate([info["primary_rays"], info["shadow_rays_buf"]], axis=0)
        return rays, info

    def algorithmic_bytes(self, passes):
        """sum over all rays of 64*N_node + 88*N_tri + 48 + 32 (SURVEY.md §8d / BASELINE.md §3)."""
        tot = dict(rays=0, n_node=0, n_tri=0, shadow=0, shadow_n_node=0, shadow_n_tri=0)
        for k in passes:
            _, _, info = self.bvh.render_pass(self.frame, W, H, rng_mode=1, pass_index=k, shader=1, light=LIGHT,
                                              nthreads=self.cores)
            tot["rays"] += info["trace_calls"] + info["shadow_rays"]
            tot["shadow"] += info["shadow_rays"]
            for key in ("n_node", "n_tri", "shadow_n_node", "shadow_n_tri"):
                tot[key] += info[key]
        tot["bytes"] = 64 * tot["n_node"] + 88 * tot["n_tri"] + 80 * tot["rays"]
        tot["shadow_bytes"] = 64 * tot["shadow_n_node"] + 88 * tot["shadow_n_tri"] + 80 * tot["shadow"]
        tot["camera_bytes"] = tot["bytes"] - tot["shadow_bytes"]
        return tot

    def trace_seconds(self, rays, repeat=1):
        """Reference Scene::Trace over the ray buffer, OpenMP schedule(dynamic,1) over 1920-ray rows."""
        if self.ref is not None:
            return self.ref.trace(rays, row=W, nthreads=self.cores, repeat=repeat)["seconds"], "reference"
        best = 1e30
        for _ in range(repeat):
            best = min(best, self.bvh.trace(rays, row=W, nthreads=self.cores)["seconds"])
        return best, "port"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    v, f = build_inputs()
    cpu = CpuSide(v, f)
    secs, nrays = [], 0
    kind = "reference"
    for i in range(args.warmup + args.steps):
        rays, _ = cpu.pass_rays(i % SPP)
        s, kind = cpu.trace_seconds(rays)
        if i >= args.warmup:
            secs.append(s)
            nrays += len(rays)
    total = float(sum(secs))
    val = nrays / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cpu.cores, "kind": kind,
                         "sample": "each step = the exact ray set of ONE of the 16 passes (2.07 M jittered camera rays + "
                                   "~0.72 M shadow rays, closest-hit Scene::Trace for both), OpenMP schedule(dynamic,1) "
                                   "over 1920-ray rows on all host threads"},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle legs (roofline bytes + cpu_baseline)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import mallie_b200 as M
    from mallie_b200 import tiles

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: one JSON line only
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    v, f = build_inputs()
    torch.zeros(1, device="cuda")
    torch.cuda.synchronize()                       # context up before the set-up is timed
    t0 = time.perf_counter()
    # BVHAccel::Build + traversal layout on the device (mb200_scene_build); outside the timed region, reported as
    # scene_build_upload_s.  The tree equals the host builder's and the reference's (tests/test_gpu_build.py).
    sc = M.Scene.build(v, f, device=local_rank, want_bvh=False)
    build_s = time.perf_counter() - t0
    frame = M.camera_frame(EYE, LOOKAT, width=W, height=H)
    stream = torch.cuda.ExternalStream(sc.stream(), device=torch.device("cuda", local_rank))
    L = M.capi.lib()
    C = M.capi.C

    bands = (BAND_ROWS, world, rank) if world > 1 else None
    params = sc.render_params(frame, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=LIGHT, pass_index=0,
                              bands=bands, compact=world > 1)
    rows_local = sc.band_local_rows(params) if world > 1 else H
    d_img = torch.zeros((rows_local, W, 3), dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros((rows_local, W), dtype=torch.int32, device="cuda")
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    gather = tiles.FramebufferGather(W, H, BAND_ROWS, world, rank, torch.device("cuda", local_rank)) if world > 1 else None

    def render_device():
        M.capi.check(L.mb200_render_frame(sc.h, C.byref(params), SPP, C.c_void_p(d_img.data_ptr()),
                                          C.c_void_p(d_cnt.data_ptr()), None))

    def step_device(ev=None):
        """One frame, everything on the device, on the scene's stream."""
        with torch.cuda.stream(stream):
            render_device()
            if ev is not None:
                ev.record(stream)
            if gather is not None:
                return gather(d_img)
        return d_img

    # exact ray counts of a frame (deterministic: the same 16 passes every step)
    _, _, st = sc.render_frame(params, SPP, d_img.data_ptr(), d_cnt.data_ptr(), stats=True)
    rays_local = st["primary_rays"] + st["shadow_rays"]
    rays_t = torch.tensor([rays_local, st["shadow_rays"]], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(rays_t)
    rays_frame, shadow_frame = int(rays_t[0]), int(rays_t[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ value: device-resident
    sampler = ClockSampler(local_rank) if rank == 0 else None   # rank 0 reports; one nvidia-smi per job
    for _ in range(args.warmup):
        step_device()
    barrier()
    sc.timing(True)                        # CUDA events around every kernel the scene launches
    launches0 = M.capi.launches_issued()
    evs = []
    barrier()
    if sampler:
        sampler.mark_start()
    for _ in range(args.steps):
        e0, em, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        with torch.cuda.stream(stream):
            flush.zero_()
            e0.record(stream)
        step_device(em)
        with torch.cuda.stream(stream):
            e1.record(stream)
        evs.append((e0, em, e1))
    barrier()
    clocks = None
    if sampler:
        sampler.mark_end()
        clocks = sampler.stop()
    launches = M.capi.launches_issued() - launches0
    kt = sc.kernel_times()
    sc.timing(False)
    step_ms = [a.elapsed_time(c) for a, _, c in evs]
    kern_ms = [a.elapsed_time(b) for a, b, _ in evs]
    total_ms = torch.tensor([sum(step_ms), float(np.mean(kern_ms)), kt["camera_trace_ms"], kt["shadow_trace_ms"],
                             kt["shade_ms"], kt["resolve_ms"], kt["trace_union_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms_max, kern_ms_max = float(total_ms[0]), float(total_ms[1])
    # totals over the timed steps; per class the UNION of the launches' spans (a frame's batches alternate between
    # two streams so that one launch's drain phase overlaps the next launch), trace_all = union over both classes
    cam_ms, shd_ms, shade_ms, resolve_ms, trace_all_ms = (float(x) for x in total_ms[2:7])
    cam_n, shd_n = int(kt["camera_trace_launches"]), int(kt["shadow_trace_launches"])
    value = rays_frame * args.steps / (total_ms_max * 1e-3) / 1e6

    # ------------------------------------------------------------------ e2e: host buffers through the C ABI
    h_img = torch.zeros((H, W, 3), dtype=torch.float32).pin_memory()
    h_cnt = torch.zeros((H, W), dtype=torch.int32).pin_memory()
    full_params = sc.render_params(frame, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=LIGHT, pass_index=0)

    def step_e2e():
        if world == 1:
            # the call a Mallie host makes: params in (by value), host framebuffer + count out
            M.capi.check(L.mb200_render_frame(sc.h, C.byref(full_params), SPP, C.c_void_p(h_img.data_ptr()),
                                              C.c_void_p(h_cnt.data_ptr()), None))
        else:
            full = step_device()
            with torch.cuda.stream(stream):
                if rank == 0:
                    h_img.copy_(full, non_blocking=True)
                    h_cnt.fill_(SPP)
            stream.synchronize()

    for _ in range(2):
        step_e2e()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        barrier()
        t0 = time.perf_counter()
        step_e2e()
        e2e_s += time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = rays_frame * args.steps / float(e2e_t[0]) / 1e6
    d2h = H * W * 3 * 4 + H * W * 4
    h2d = C.sizeof(M.capi.RenderParams)

    # ------------------------------------------------------------------ oracle legs (rank 0)
    roofline, cpu_baseline = None, None
    peak, peak_src = hbm_peak()
    if rank == 0 and not args.no_cpu:
        cpu = CpuSide(v, f)
        alg = cpu.algorithmic_bytes(range(SPP))
        assert alg["rays"] == rays_frame, (alg["rays"], rays_frame)
        # every rank renders 1/world of the row bands; per-launch bytes = the frame's bytes / launches per frame
        per_frame = {"camera": alg["camera_bytes"] / world, "shadow": alg["shadow_bytes"] / world}
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tp) and world == 1:
            with open(tp) as fp:
                traffic = json.load(fp)
        kernels = []
        for name, ms, n, inst in (("camera", cam_ms, cam_n, "k_trace_sm<IOCamera, closest-hit> (raygen fused)"),
                                  ("shadow", shd_ms, shd_n, "k_trace_sm<IOQueueShadow, any-hit>")):
            if n == 0:
                continue
            b = per_frame[name] * args.steps / n
            a = b / (ms / n * 1e-3) / 1e9
            kernels.append({"kernel": inst, "launches_per_step": n / args.steps, "ms_per_launch": ms / n,
                            "algorithmic_bytes_per_launch": b, "achieved": a, "frac": a / peak,
                            "traffic": traffic.get(name + "_trace_dram_bytes_per_launch")})
        trace_ms, trace_n = trace_all_ms, cam_n + shd_n
        bytes_launch = (per_frame["camera"] + per_frame["shadow"]) * args.steps / trace_n
        achieved = bytes_launch / (trace_ms / trace_n * 1e-3) / 1e9
        tr = [k["traffic"] for k in kernels]
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (sum(t * k["launches_per_step"] for t, k in zip(tr, kernels)) /
                                sum(k["launches_per_step"] for k in kernels)) if all(tr) and tr else None,
                    "peak_source": peak_src,
                    "kernel": "k_trace_sm (persistent-warp BVH traversal state machine): all launches of a step; "
                              "duration = time with at least one of them in flight / launches (launches of consecutive "
                              "batches overlap on two streams)",
                    "launches_per_step": trace_n / args.steps, "ms_per_launch": trace_ms / trace_n,
                    "algorithmic_bytes_per_launch": bytes_launch, "kernels": kernels,
                    # share of the step with a traversal launch in flight; the shade / resolve launches run inside the
                    # drain phases of the other stream's traversal launch (their event spans include that wait:
                    # ncu's serialised launch list, profiles/r1_launches_sm.csv, has trace 92 %, shade 7.5 %)
                    "step_share": {"trace": trace_ms / args.steps / (total_ms_max / args.steps)},
                    "span_ms_per_step_incl_queueing": {"shade": shade_ms / args.steps, "resolve": resolve_ms / args.steps},
                    "bytes_per_ray": alg["bytes"] / alg["rays"],
                    "nodes_per_ray": alg["n_node"] / alg["rays"], "tris_per_ray": alg["n_tri"] / alg["rays"],
                    "camera_nodes_per_ray": (alg["n_node"] - alg["shadow_n_node"]) / (alg["rays"] - alg["shadow"]),
                    "camera_tris_per_ray": (alg["n_tri"] - alg["shadow_n_tri"]) / (alg["rays"] - alg["shadow"]),
                    "shadow_nodes_per_ray": alg["shadow_n_node"] / max(1, alg["shadow"]),
                    "shadow_tris_per_ray": alg["shadow_n_tri"] / max(1, alg["shadow"])}
        if world == 1:
            sample_passes = 4
            rays = np.concatenate([cpu.pass_rays(k)[0] for k in range(sample_passes)], axis=0)
            sec, kind = cpu.trace_seconds(rays, repeat=2)
            cpu_baseline = {"value": len(rays) / sec / 1e6, "unit": "Mrays/s", "cores": cpu.cores, "kind": kind,
                            "sample": f"{sample_passes} of the {SPP} passes, all pixels: {len(rays)} rays (camera + shadow, "
                                      "closest-hit Scene::Trace), OpenMP schedule(dynamic,1) over 1920-ray rows, best of 2"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * float(e2e_t[0]) / args.steps},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "rays_per_step": rays_frame, "shadow_rays_per_step": shadow_frame, "scene_build_upload_s": build_s,
        }
        emit(line)
    # tensors that were used on the scene's stream must be released before the stream is destroyed
    # (torch's caching allocator records an event on that stream when it frees them)
    del d_img, d_cnt, flush, gather, h_img, h_cnt, evs
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sc.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
