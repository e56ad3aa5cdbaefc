#!/usr/bin/env python
"""bench.py -- Mrays/s of the Mallie render hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1m|10m] [--impl reference] [--no-cpu]

Workloads
  1m  (default; BASELINE.json configs[3], the one the north-star target is quoted on): bumpy sphere N=500 -> exactly
      1 000 000 triangles, 1920x1080, 16 spp, primary closest-hit ray + one shadow ray per primary hit.  84 MB of
      scene: L2-resident within a step.
  10m (configs[4]): N=1581 -> 9 998 244 triangles, 3840x2160, 64 spp.  0.85 GB of scene: spills the 126 MB L2, the one
      workload where the HBM roofline is a physical bound.
A "step" is one frame.  Camera eye (0,0,3) -> origin, fov 45, light (2,4,3).

* value : whole-job Mrays/s with everything resident in HBM (the assembled framebuffer stays on the device).
          N > 1: one process per GPU, image rows interleaved over the ranks in bands of 4 scanlines (strong scaling:
          the frame is fixed), every rank renders its bands into a compact buffer and mb200_render_frame_gathered
          assembles the frame on every rank: one kernel of the library stores the rows into every rank's frame buffer
          over peer memory (or, MB200_GATHER=nccl / when buffers cannot be mapped: ncclAllGather + row placement).
* e2e   : the same frame through the C-ABI call a Mallie host makes with a pinned HOST framebuffer
          (mb200_render_frame at N = 1, mb200_render_frame_gathered at N > 1 with the host buffer on rank 0); the
          device->host copy of the framebuffer is inside the timed region (at N = 1 the library copies the rows of the
          frame's first batch while the second one is still traced).
* parity: outside the timed region, at every N: a 64-bit digest of the final frame on every rank, compared with the oracle's
          frame (all 16 passes for 1m; the first pass for 10m, and says so) -- the run FAILS if they differ.
* roofline : the traversal kernel k_trace_sm (closest-hit launches over camera rays + any-hit launches over shadow rays).
    hbm   achieved = algorithmic bytes per launch (64 B/node popped + 88 B/triangle tested + 48 B ray + 32 B hit record,
          counted by the CPU oracle in reference traversal order for the exact ray set; a shadow ray billed as the
          closest-hit Traverse that defines its oracle) / average launch duration, timed with CUDA events around every
          launch on its stream inside the timed region (mb200_scene_timing), against the measured HBM copy bandwidth
          (MEASURED_PEAKS.json).  `shadow_billing` gives the same figure with the shadow rays billed at what the
          any-hit walk really visited (the GPU's own counters).  On the 1m scene this fraction exceeds 1 because the
          scene is L2-resident (traffic = ncu DRAM bytes per launch, ~1 % of the algorithmic bytes): it says the
          kernel is not HBM-bound, nothing more.  The two fractions that bind:
    fp64  the reference's FP64 operations (19 per box test: 6 sub, 6 mul, 7 compares; 59 per triangle test: 27 mul,
          24 add/sub, 1 div, 7 compares) x the oracle's counts / launch time, against the FP64-pipe lane-op rate
          measured in this run (mb200_probe_peaks: independent DADD/DMUL chains).
    l2    ncu lts__t_bytes per launch (profiles/, provenance in roofline.l2.source) / launch time, against the L2 read
          bandwidth measured in this run.
* cpu_baseline : the unmodified reference (oracle/_ref) tracing a sample of the same ray set on the host cores.
* --impl reference : times the reference's own OpenMP CPU path on the same workload (rank 0 only).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EYE, LOOKAT, LIGHT = (0.0, 0.0, 3.0), (0.0, 0.0, 0.0), (2.0, 4.0, 3.0)
BAND_ROWS = 4          # one tile row per band: the finest interleave (rank load differs by < 1 band in ~22)
L2_FLUSH_BYTES = 512 << 20
WORKLOADS = {
    "1m": dict(sphere_n=500, W=1920, H=1080, spp=16, triangles=1000000, vertices=501501, scene_mb=84,
               config="BASELINE.json configs[3]", oracle_passes=16),
    "10m": dict(sphere_n=1581, W=3840, H=2160, spp=64, triangles=9998244, vertices=5003866, scene_mb=850,
                config="BASELINE.json configs[4]", oracle_passes=1),
}
FP64_OPS_PER_BOX, FP64_OPS_PER_TRI = 19, 59


def metric_name(wl):
    return f"Mrays/s primary+shadow at {wl['W']}x{wl['H']}"


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


EXCHANGE = {1: "k_exchange_rows: every rank stores its rows straight into every rank's frame buffer over peer memory "
                 "(cudaIpc-mapped through the communicator), flag barrier in the same kernel",
            0: "one ncclAllGather of the framebuffer per frame + the library's row-placement kernel"}


def config_dict(wl, name, n_gpus, exchange=None):
    return {
        "workload": f"{name}: bumpy-sphere N={wl['sphere_n']} ({wl['triangles']:,} triangles, {wl['vertices']:,} vertices), "
                    f"{wl['W']}x{wl['H']}, {wl['spp']} spp, primary closest-hit + 1 shadow ray per hit, eye (0,0,3) lookat "
                    f"(0,0,0) fov 45, light (2,4,3) [{wl['config']}]",
        "triangles": wl["triangles"], "resolution": [wl["W"], wl["H"]], "spp": wl["spp"], "shader": "primary+shadow",
        "parallelism": "1 GPU" if n_gpus == 1 else f"image rows in {BAND_ROWS}-scanline bands interleaved over "
                                                    f"{n_gpus} GPUs (one process each), full scene replica per GPU, "
                                                    f"frame assembled on every rank by mb200_render_frame_gathered (C ABI): "
                                                    f"{EXCHANGE.get(exchange, EXCHANGE[0])}",
        "l2": f"L2 flushed between timed steps (512 MiB memset); the scene ({wl['scene_mb']} MB) " +
              ("is L2-resident within a step" if wl["scene_mb"] < 126 else "does not fit the 126 MB L2"),
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


def frame_hash(a):
    """64-bit digest of a frame's bytes (identity check between ranks, host / device copies and the oracle's frame)."""
    return hashlib.blake2b(np.ascontiguousarray(a).tobytes(), digest_size=8).hexdigest()


def build_inputs(wl):
    from mallie_b200.procedural import bumpy_sphere
    return bumpy_sphere(wl["sphere_n"])


# ----------------------------------------------------------------------------------------------------
# CPU side (oracle / reference): test infrastructure, used here only as checker + baseline
# ----------------------------------------------------------------------------------------------------
class CpuSide:
    def __init__(self, v, f, wl):
        from oracle import orabind as O
        from oracle import refbind as R
        self.O, self.R, self.wl = O, R, wl
        self.mesh = O.Mesh(v, f)
        self.bvh = O.BVH.build(self.mesh)
        self.frame = O.camera_frame(EYE, LOOKAT, width=wl["W"], height=wl["H"])
        self.ref = None
        if R.available():
            self.ref = R.RefScene.from_arrays(v, f)
            self.ref.build()
        # all host threads, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def pass_rays(self, k):
        """Exact ray set of pass k (jittered camera rays + shadow rays)."""
        _, _, info = self.bvh.render_pass(self.frame, self.wl["W"], self.wl["H"], rng_mode=1, pass_index=k, shader=1,
                                          light=LIGHT, emit_rays=True, nthreads=self.cores)
        rays = np.concatenate([info["primary_rays"], info["shadow_rays_buf"]], axis=0)
        return rays, info

    def render(self, passes):
        """The oracle's frame over `passes` (AccumImage order: float += float, pass after pass) and its counters:
        sum over all rays of 64*N_node + 88*N_tri + 48 + 32 (SURVEY.md §8d / BASELINE.md §3)."""
        W, H = self.wl["W"], self.wl["H"]
        tot = dict(rays=0, n_node=0, n_tri=0, shadow=0, shadow_n_node=0, shadow_n_tri=0, passes=len(passes))
        image = np.zeros((H, W, 3), np.float32)
        for k in passes:
            img, _, info = self.bvh.render_pass(self.frame, W, H, rng_mode=1, pass_index=k, shader=1, light=LIGHT,
                                                nthreads=self.cores)
            image += img
            tot["rays"] += info["trace_calls"] + info["shadow_rays"]
            tot["shadow"] += info["shadow_rays"]
            for key in ("n_node", "n_tri", "shadow_n_node", "shadow_n_tri"):
                tot[key] += info[key]
        tot["bytes"] = 64 * tot["n_node"] + 88 * tot["n_tri"] + 80 * tot["rays"]
        tot["shadow_bytes"] = 64 * tot["shadow_n_node"] + 88 * tot["shadow_n_tri"] + 80 * tot["shadow"]
        tot["camera_bytes"] = tot["bytes"] - tot["shadow_bytes"]
        tot["image_hash"] = frame_hash(image)
        return tot

    def trace_seconds(self, rays, repeat=1):
        """Reference Scene::Trace over the ray buffer, OpenMP schedule(dynamic,1) over image-width rows."""
        if self.ref is not None:
            return self.ref.trace(rays, row=self.wl["W"], nthreads=self.cores, repeat=repeat)["seconds"], "reference"
        best = 1e30
        for _ in range(repeat):
            best = min(best, self.bvh.trace(rays, row=self.wl["W"], nthreads=self.cores)["seconds"])
        return best, "port"


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    v, f = build_inputs(wl)
    cpu = CpuSide(v, f, wl)
    secs, nrays = [], 0
    kind = "reference"
    for i in range(args.warmup + args.steps):
        rays, _ = cpu.pass_rays(i % wl["spp"])
        s, kind = cpu.trace_seconds(rays)
        if i >= args.warmup:
            secs.append(s)
            nrays += len(rays)
    total = float(sum(secs))
    val = nrays / total / 1e6
    line = {
        "impl": "reference", "metric": metric_name(wl), "value": val, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(wl, args.workload, args.gpus),
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cpu.cores, "kind": kind,
                         "sample": f"each step = the exact ray set of ONE of the {wl['spp']} passes (jittered camera rays + "
                                   "their shadow rays, closest-hit Scene::Trace for both), OpenMP schedule(dynamic,1) "
                                   "over image-width rows on all host threads"},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


def ncu_traffic(name):
    """DRAM / L2 bytes per launch from the committed ncu captures of this workload (profiles/r2_traffic.json), with
    their provenance; {} when there is none."""
    tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(tp):
        return {}
    with open(tp) as fp:
        return json.load(fp).get(name, {})


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1m", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle legs (frame parity vs the oracle, roofline bytes, cpu_baseline)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    W, H, SPP = wl["W"], wl["H"], wl["spp"]
    quiet_stdout()
    if args.impl == "reference":
        return run_reference_arm(args, wl)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import mallie_b200 as M

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

    v, f = build_inputs(wl)
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

    params = sc.render_params(frame, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=LIGHT, pass_index=0)
    d_img = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")      # the assembled frame (every rank)
    d_cnt = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    comm = None
    if world > 1:
        # the library's own NCCL communicator (C ABI): rank 0's unique id travels over the torch process group
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(M.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm = M.Comm(sc, world, rank, bytes(uid.cpu().numpy().tobytes()))

    def render(image_ptr, count_ptr, stats=False):
        """One frame through the C ABI: the whole frame at N = 1, this rank's bands + the NCCL gather at N > 1."""
        if comm is None:
            st = M.capi.RenderStats()
            M.capi.check(L.mb200_render_frame(sc.h, C.byref(params), SPP, image_ptr, count_ptr, C.byref(st) if stats else None))
            return st.as_dict() if stats else None
        st = M.capi.RenderStats()
        M.capi.check(L.mb200_render_frame_gathered(comm.h, C.byref(params), SPP, BAND_ROWS, image_ptr, count_ptr,
                                                   C.byref(st) if stats else None))
        return st.as_dict() if stats else None

    dev_img, dev_cnt = C.c_void_p(d_img.data_ptr()), C.c_void_p(d_cnt.data_ptr())

    # exact ray counts and traversal counters of a frame (deterministic: the same passes every step)
    st = render(dev_img, dev_cnt, stats=True)
    keys = ("primary_rays", "shadow_rays", "camera_nodes_tested", "camera_tris_tested", "shadow_nodes_tested", "shadow_tris_tested")
    cnt_t = torch.tensor([st[k] for k in keys], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt_t)
    gpu_counts = dict(zip(keys, (int(x) for x in cnt_t)))
    rays_frame, shadow_frame = gpu_counts["primary_rays"] + gpu_counts["shadow_rays"], gpu_counts["shadow_rays"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ value: device-resident
    sampler = ClockSampler(local_rank) if rank == 0 else None   # rank 0 reports; one nvidia-smi per job
    for _ in range(args.warmup):
        render(dev_img, dev_cnt)
    barrier()
    sc.timing(True)                        # CUDA events around every kernel the scene launches
    launches0 = M.capi.launches_issued()
    evs = []
    barrier()
    if sampler:
        sampler.mark_start()
    for _ in range(args.steps):
        e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        with torch.cuda.stream(stream):
            flush.zero_()
            e0.record(stream)
        render(dev_img, dev_cnt)
        with torch.cuda.stream(stream):
            e1.record(stream)
        evs.append((e0, e1))
    barrier()
    clocks = None
    if sampler:
        sampler.mark_end()
        clocks = sampler.stop()
    launches = M.capi.launches_issued() - launches0
    kt = sc.kernel_times()
    sc.timing(False)
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms), kt["camera_trace_ms"], kt["shadow_trace_ms"], kt["shade_ms"], kt["resolve_ms"],
                             kt["trace_union_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms_max = float(total_ms[0])
    # totals over the timed steps; per class the UNION of the launches' spans (a frame's batches alternate between
    # two streams so that one launch's drain phase overlaps the next launch), trace_all = union over both classes
    cam_ms, shd_ms, shade_ms, resolve_ms, trace_all_ms = (float(x) for x in total_ms[1:6])
    cam_n, shd_n = int(kt["camera_trace_launches"]), int(kt["shadow_trace_launches"])
    value = rays_frame * args.steps / (total_ms_max * 1e-3) / 1e6

    # ------------------------------------------------------------------ parity: the frame of the timed loop, every rank
    frame_fnv = frame_hash(d_img.cpu().numpy())
    counts_ok = bool((d_cnt == SPP).all().item())
    fnv_all = [frame_fnv]
    if world > 1:
        fnv_all = [None] * world
        dist.all_gather_object(fnv_all, frame_fnv)
    ranks_equal = len(set(fnv_all)) == 1

    # ------------------------------------------------------------------ e2e: host framebuffer through the C ABI
    h_img = torch.zeros((H, W, 3), dtype=torch.float32).pin_memory()
    h_cnt = torch.zeros((H, W), dtype=torch.int32).pin_memory()
    host_img = C.c_void_p(h_img.data_ptr()) if rank == 0 else None      # N > 1: the frame is delivered to rank 0's host
    host_cnt = C.c_void_p(h_cnt.data_ptr()) if rank == 0 else None

    def step_e2e():
        if comm is None:
            render(host_img, host_cnt)             # the call a Mallie host makes: params in (by value), host framebuffer out
        else:
            render(host_img, host_cnt)
            sc.synchronize()

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
    d2h = H * W * 3 * 4          # the float frame; count of a fresh whole frame is the constant SPP, written by the host
    h2d = C.sizeof(M.capi.RenderParams)
    host_fnv = frame_hash(h_img.numpy()) if rank == 0 else None

    # ------------------------------------------------------------------ measured ceilings of this GPU (rank 0)
    peaks = M.capi.probe_peaks(local_rank) if rank == 0 else None

    # ------------------------------------------------------------------ oracle legs (rank 0)
    roofline, cpu_baseline = None, None
    parity = {"frame_hash": frame_fnv, "ranks": world, "ranks_equal": ranks_equal, "host_frame_hash": host_fnv,
              "count_is_spp_everywhere": counts_ok, "oracle_hash": None, "oracle_checked": "skipped (--no-cpu)"}
    peak, peak_src = hbm_peak()
    if rank == 0 and not args.no_cpu:
        cpu = CpuSide(v, f, wl)
        npass = wl["oracle_passes"]
        alg = cpu.render(range(npass))
        if npass == SPP:
            parity.update(oracle_hash=alg["image_hash"], oracle_checked=f"all {SPP} passes: the frame of the timed loop")
            assert alg["rays"] == rays_frame, (alg["rays"], rays_frame)
            gpu_fnv_for_oracle = frame_fnv
        else:
            # bounded oracle leg (the 10m frame is 714 M rays): the GPU re-renders the first `npass` passes, those are compared
            one = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
            M.capi.check(L.mb200_render_frame(sc.h, C.byref(params), npass, C.c_void_p(one.data_ptr()), dev_cnt, None))
            sc.synchronize()
            gpu_fnv_for_oracle = frame_hash(one.cpu().numpy())
            parity.update(oracle_hash=alg["image_hash"], gpu_hash_same_passes=gpu_fnv_for_oracle,
                          oracle_checked=f"first {npass} of {SPP} passes (bounded CPU leg), whole image")
            del one
        parity["equal_to_oracle"] = gpu_fnv_for_oracle == alg["image_hash"]
        scale = SPP / npass                                    # per-frame figures from the oracle's passes
        # every rank renders 1/world of the row bands; per-launch bytes = the frame's bytes / launches per frame
        per_frame = {"camera": alg["camera_bytes"] * scale / world, "shadow": alg["shadow_bytes"] * scale / world}
        ops_frame = {"camera": (FP64_OPS_PER_BOX * (alg["n_node"] - alg["shadow_n_node"]) +
                                FP64_OPS_PER_TRI * (alg["n_tri"] - alg["shadow_n_tri"])) * scale / world,
                     "shadow": (FP64_OPS_PER_BOX * alg["shadow_n_node"] + FP64_OPS_PER_TRI * alg["shadow_n_tri"]) * scale / world}
        anyhit_bytes = (64 * gpu_counts["shadow_nodes_tested"] + 88 * gpu_counts["shadow_tris_tested"] + 80 * shadow_frame) / world
        anyhit_ops = (FP64_OPS_PER_BOX * gpu_counts["shadow_nodes_tested"] + FP64_OPS_PER_TRI * gpu_counts["shadow_tris_tested"]) / world
        traffic = ncu_traffic(args.workload) if world == 1 else {}
        kernels = []
        for name, ms, n, inst in (("camera", cam_ms, cam_n, "k_trace_sm<IOCameraT, closest-hit> (raygen fused)"),
                                  ("shadow", shd_ms, shd_n, "k_trace_sm<IOQueueShadow, any-hit>")):
            if n == 0:
                continue
            sec = ms / n * 1e-3
            b = per_frame[name] * args.steps / n
            a = b / sec / 1e9
            k = {"kernel": inst, "launches_per_step": n / args.steps, "ms_per_launch": ms / n,
                 "algorithmic_bytes_per_launch": b, "achieved": a, "frac": a / peak,
                 "traffic": traffic.get(name + "_dram_bytes_per_launch"),
                 "fp64_frac": ops_frame[name] * args.steps / n / sec / peaks["fp64_lane_ops_per_s"]}
            l2b = traffic.get(name + "_lts_bytes_per_launch")
            if l2b:
                k["l2_frac"] = l2b / sec / peaks["l2_read_bytes_per_s"]
            kernels.append(k)
        trace_ms, trace_n = trace_all_ms, cam_n + shd_n
        sec_launch = trace_ms / trace_n * 1e-3
        bytes_launch = (per_frame["camera"] + per_frame["shadow"]) * args.steps / trace_n
        achieved = bytes_launch / sec_launch / 1e9
        bytes_launch_anyhit = (per_frame["camera"] + anyhit_bytes) * args.steps / trace_n
        ops_launch = (ops_frame["camera"] + ops_frame["shadow"]) * args.steps / trace_n
        ops_launch_anyhit = (ops_frame["camera"] + anyhit_ops) * args.steps / trace_n
        tr = [k["traffic"] for k in kernels]
        lts = [traffic.get(nm + "_lts_bytes_per_launch") for nm in ("camera", "shadow")]
        l2_leg = None
        if all(lts) and cam_n and shd_n:
            l2_bytes_launch = (lts[0] * cam_n + lts[1] * shd_n) / trace_n
            l2_leg = {"achieved": l2_bytes_launch / sec_launch / 1e9, "peak": peaks["l2_read_bytes_per_s"] / 1e9, "unit": "GB/s",
                      "frac": l2_bytes_launch / sec_launch / peaks["l2_read_bytes_per_s"],
                      "lts_bytes_per_launch": l2_bytes_launch, "source": traffic.get("source")}
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (sum(t * k["launches_per_step"] for t, k in zip(tr, kernels)) /
                                sum(k["launches_per_step"] for k in kernels)) if tr and all(tr) else None,
                    "traffic_source": traffic.get("source"),
                    "peak_source": peak_src,
                    "kernel": "k_trace_sm (persistent-warp BVH traversal state machine): all launches of a step; "
                              "duration = time with at least one of them in flight / launches (launches of consecutive "
                              "batches overlap on two streams)",
                    "launches_per_step": trace_n / args.steps, "ms_per_launch": trace_ms / trace_n,
                    "algorithmic_bytes_per_launch": bytes_launch, "kernels": kernels,
                    "shadow_billing": {
                        "closest_hit": {"what": "a shadow ray = the closest-hit Traverse that defines its oracle (the headline figure)",
                                        "nodes_per_shadow_ray": alg["shadow_n_node"] / max(1, alg["shadow"]),
                                        "tris_per_shadow_ray": alg["shadow_n_tri"] / max(1, alg["shadow"]),
                                        "achieved": achieved, "frac": achieved / peak},
                        "any_hit_gpu_counters": {"what": "a shadow ray = what the any-hit walk visited (mb200_render_stats counters of this run)",
                                                 "nodes_per_shadow_ray": gpu_counts["shadow_nodes_tested"] / max(1, shadow_frame),
                                                 "tris_per_shadow_ray": gpu_counts["shadow_tris_tested"] / max(1, shadow_frame),
                                                 "achieved": bytes_launch_anyhit / sec_launch / 1e9,
                                                 "frac": bytes_launch_anyhit / sec_launch / 1e9 / peak}},
                    "fp64": {"achieved": ops_launch / sec_launch / 1e12, "peak": peaks["fp64_lane_ops_per_s"] / 1e12,
                             "unit": "T lane-ops/s", "frac": ops_launch / sec_launch / peaks["fp64_lane_ops_per_s"],
                             "frac_any_hit_billing": ops_launch_anyhit / sec_launch / peaks["fp64_lane_ops_per_s"],
                             "ops_per_box_test": FP64_OPS_PER_BOX, "ops_per_triangle_test": FP64_OPS_PER_TRI,
                             "how": "reference FP64 operations x oracle counts / launch time vs DADD/DMUL lane-op rate "
                                    "measured in this run (mb200_probe_peaks)"},
                    "l2": l2_leg,
                    "measured_peaks": peaks,
                    # share of the step with a traversal launch in flight; the shade / resolve launches run inside the
                    # drain phases of the other stream's traversal launch (their event spans include that wait)
                    "step_share": {"trace": trace_ms / args.steps / (total_ms_max / args.steps)},
                    "span_ms_per_step_incl_queueing": {"shade": shade_ms / args.steps, "resolve": resolve_ms / args.steps},
                    "bytes_per_ray": alg["bytes"] / alg["rays"],
                    "nodes_per_ray": alg["n_node"] / alg["rays"], "tris_per_ray": alg["n_tri"] / alg["rays"],
                    "camera_nodes_per_ray": (alg["n_node"] - alg["shadow_n_node"]) / (alg["rays"] - alg["shadow"]),
                    "camera_tris_per_ray": (alg["n_tri"] - alg["shadow_n_tri"]) / (alg["rays"] - alg["shadow"]),
                    "shadow_nodes_per_ray": alg["shadow_n_node"] / max(1, alg["shadow"]),
                    "shadow_tris_per_ray": alg["shadow_n_tri"] / max(1, alg["shadow"])}
        if world == 1:
            sample_passes = 4 if args.workload == "1m" else 1
            rays = np.concatenate([cpu.pass_rays(k)[0] for k in range(sample_passes)], axis=0)
            sec, kind = cpu.trace_seconds(rays, repeat=2)
            cpu_baseline = {"value": len(rays) / sec / 1e6, "unit": "Mrays/s", "cores": cpu.cores, "kind": kind,
                            "sample": f"{sample_passes} of the {SPP} passes, all pixels: {len(rays)} rays (camera + shadow, "
                                      f"closest-hit Scene::Trace), OpenMP schedule(dynamic,1) over {W}-ray rows, best of 2"}

    ok = ranks_equal and counts_ok and parity.get("equal_to_oracle", True) and (host_fnv in (None, frame_fnv))
    if rank == 0:
        line = {
            "metric": metric_name(wl), "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(wl, args.workload, world, comm.exchange_path() if comm is not None else None), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * float(e2e_t[0]) / args.steps},
            "gpu_launches": launches, "parity": parity, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "rays_per_step": rays_frame, "shadow_rays_per_step": shadow_frame, "scene_build_upload_s": build_s,
        }
        emit(line)
    # tensors that were used on the scene's stream must be released before the stream is destroyed
    # (torch's caching allocator records an event on that stream when it frees them)
    del d_img, d_cnt, flush, h_img, h_cnt, evs
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sc.close()
    if not ok:
        print(f"bench.py: PARITY FAILURE: {parity}", file=sys.stderr)
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
