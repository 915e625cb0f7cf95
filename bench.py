#!/usr/bin/env python
"""bench.py — headline benchmark of the FLaME variational-depth hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Workload (BASELINE.json configs[3], the one `metric` is quoted on): 1920x1080 frames, 10 000
Delaunay vertices each, 200 primal-dual iterations + dense inverse-depth projection.
A "step" is one pass of the hot path over one batch of B independent synthetic frames per GPU:
    graph_reset (fresh-vertex rule)  ->  200 x step()  ->  interpolateMesh
`value`  = frames/s with inputs resident in HBM (CUDA events on the library's stream, max over ranks).
`e2e`    = the same metric through the public C ABI with HOST buffers: per frame, upload of the
           graph + triangles (H2D), PD, projection, download of the dense map + vertex idepths (D2H).
`roofline` is for the dominant kernel (the resident PD solver) — see DESIGN.md for the byte model.
Multi-GPU: frames shard by rank (weak scaling), no data-path collective; torch.distributed (NCCL)
is only used for the barrier and the max-over-ranks reduction.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = "D"
METRIC = "frames/sec at 1080p (10k vertices, 200 PD iters + dense projection)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _frames(n, first=0):
    from flame_b200 import synth
    W, H, V, iters = synth.CONFIGS[CFG]
    out = []
    for f in range(first, first + n):
        g = synth.config_graph(CFG, frame=f)
        out.append(g)
    return out, (W, H, V, iters)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # pragma: no cover
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def _cpu_frame(pyoracle, synth, g, iters):
    """One frame of the hot path on one CPU thread (the reference runs step() on exactly one thread,
    flame.cc:100-111): edge weights, 200 x step, interpolateMesh."""
    alpha, beta = pyoracle.edge_weights(g["edges"], g["pos"])
    s = synth.init_state(g, alpha, beta)
    if pyoracle.have_ref():  # the reference's own sources compiled verbatim (oracle/_ref)
        pyoracle.ref_pd_run(s, pyoracle.pd_default_params(), iters)
        pyoracle.interpolate_mesh(g["tris"], g["pos"], s["x"], g["H"], g["W"], use_ref=True)
    else:                    # the oracle's C restatement
        pyoracle.pd_run(s, pyoracle.pd_default_params(), iters)
        pyoracle.interpolate_mesh(g["tris"], g["pos"], s["x"], g["H"], g["W"])
    return s


def _cpu_kind(pyoracle):
    return "reference" if pyoracle.have_ref() else "port"


_CPU_DESC = {"reference": "reference sources compiled verbatim against oracle/shim (oracle/_ref)",
             "port": "oracle C port of the reference CPU path"}


def cpu_baseline(budget_s=12.0):
    from flame_b200 import synth
    from oracle import pyoracle
    pyoracle.build()
    frames, (W, H, V, iters) = _frames(1)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < budget_s:
        _cpu_frame(pyoracle, synth, frames[0], iters)
        n += 1
    dt = time.perf_counter() - t0
    out = {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": _cpu_kind(pyoracle),
           "sample": "%d frames of config %s (%dx%d, V=%d, %d PD iters + dense projection), %s, "
                     "1 thread (the reference runs step() on one thread), %.1f s"
                     % (n, CFG, W, H, V, iters, _CPU_DESC[_cpu_kind(pyoracle)], dt)}
    try:  # the epipolar search of the `search` side measurement on one CPU thread (OpenMP is inert in the
        # reference build): same workload, config C
        Wc, Hc, _, _ = synth.CONFIGS["C"]
        K, Kinv, sframes = synth.make_stereo_pair(Wc, Hc)
        feats0 = synth.make_features(synth.config_graph("C")["pos"], Wc, Hc)
        fn = pyoracle.ref_update_feature_idepths if pyoracle.have_ref() else pyoracle.update_feature_idepths
        tp = pyoracle.track_default_params()
        ts = []
        for _ in range(3):
            f = feats0.copy()
            t0 = time.perf_counter()
            fn(tp, K, Kinv, Wc, Hc, synth.PAD, sframes, 1, 0, f)
            ts.append(time.perf_counter() - t0)
        out["search"] = {"ms_per_call": 1e3 * min(ts), "features_per_s": len(feats0) / min(ts), "cores": 1,
                         "workload": "config C: %dx%d, %d features" % (Wc, Hc, len(feats0))}
    except Exception as e:  # the headline baseline above does not depend on it
        out["search"] = {"unavailable": repr(e)}
    return out


def _ref_worker(conn, n_frames_max, iters):
    """One worker PROCESS = one Flame instance on one core: runs whole frames until told to stop."""
    from flame_b200 import synth
    from oracle import pyoracle
    frames, _ = _frames(1)
    g = frames[0]
    conn.send("ready")
    while True:
        n = conn.recv()
        if n <= 0:
            break
        for _ in range(n):
            _cpu_frame(pyoracle, synth, g, iters)
        conn.send(n)


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores: independent frames in
    parallel, one single-threaded worker process per core (each Flame instance runs step() on one
    thread, flame.cc:100-111). A step = one frame per core."""
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import pyoracle
    pyoracle.build()
    cores = os.cpu_count() or 1
    W, H, V, iters = __import__("flame_b200.synth", fromlist=["CONFIGS"]).CONFIGS[CFG]
    ctx = mp.get_context("fork")
    workers = []
    for _ in range(cores):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(b, 0, iters), daemon=True)
        p.start()
        workers.append((p, a))
    for _, a in workers:
        a.recv()
    per_step = cores  # one frame per core per step

    def step():
        for _, a in workers:
            a.send(1)
        for _, a in workers:
            a.recv()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    for p, a in workers:
        a.send(0)
    for p, _ in workers:
        p.join(timeout=10)
    val = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "config D: 1920x1080, V=%d, %d PD iters + dense projection" % (V, iters),
                   "frames_per_step": per_step},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": _cpu_kind(pyoracle),
                         "sample": "%d frames per step (one per core; one single-threaded worker process per "
                                   "core, as each Flame instance runs step() on one thread), %s"
                                   % (per_step, _CPU_DESC[_cpu_kind(pyoracle)])},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def _pin_to_gpu_node(index):
    """Run this rank on the CPUs of its GPU's NUMA node: the pinned host buffers of the e2e leg are then
    allocated next to the GPU's PCIe root (measured: 36 -> 39 GB/s D2H on the 2-socket bench box)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = os.cpu_count()
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = [i for i in range(n) if (words[i // 64] >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def run_ours(args, rank, world, local_rank):
    import torch
    from flame_b200 import _lib

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _pin_to_gpu_node(local_rank)

    from flame_b200 import shard
    B = args.batch
    ids = shard.frames_for_rank(rank, world, B)  # frames shard by rank: no data-path collective
    frames, (W, H, V, iters) = _frames(B, first=ids[0])
    p = _lib.default_pd_params()
    ctx = _lib.Context(local_rank, B)
    for i, g in enumerate(frames):
        ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
        ctx.mesh_bind(i, g["tris"], H, W)
    ctx.sync()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    # A frame is one 16-CTA cluster; the device keeps `wave` of them resident (7 on a B200). A step walks its
    # B frames wave by wave; the projection of a wave runs on the context's second stream and overlaps the
    # solver of the next wave (the clusters leave 36 of 148 SMs free).
    wave = args.wave if args.wave > 0 else (ctx.pd_wave(0) if args.mode != 1 else 0)
    if wave <= 0:
        wave = B
    wave = min(wave, B)
    ctx.set_async_projection(not args.no_async)
    waves = [(s0, min(wave, B - s0)) for s0 in range(0, B, wave)]

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(ev=None):
        for w, (s0, n) in enumerate(waves):
            ctx.graph_reset(s0, n)
            if ev is not None:
                ev[w][0].record(stream)
            ctx.pd_run(p, iters, s0, n, args.mode)
            if ev is not None:
                ev[w][1].record(stream)
            ctx.mesh_project(s0, n, 1.0)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: device-resident inputs ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    kev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in waves]
           for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.kernel_launches()
    e0.record(stream)
    for k in range(args.steps):
        step(kev[k])
    ctx.join()  # the last projections (second stream) belong to the timed region
    e1.record(stream)
    barrier()
    launches = _lib.kernel_launches() - launches0
    ms_total = e0.elapsed_time(e1)
    # average duration of ONE solver launch (a full wave) inside the timed region
    full = [a.elapsed_time(b) for st in kev for (a, b), (s0, n) in zip(st, waves) if n == wave]
    pd_ms = float(np.mean(full))
    clocks = sampler.stop()
    ms_total = shard.max_over_ranks(ms_total, dev)
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- single-frame latency (one cluster per frame, B = 1) ----
    lat = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.graph_reset(0, 1)
        a.record(stream)
        ctx.pd_run(p, iters, 0, 1, args.mode)
        b.record(stream)
        ctx.mesh_project(0, 1, 1.0)
        ctx.sync()
        lat.append(a.elapsed_time(b))
    lat_ms = float(np.median(lat))

    # ---- epipolar search (config C: 1280x720, 5000 features), reported as features/s ----
    search = None
    if rank == 0 and not args.no_search:
        from flame_b200 import synth
        Wc, Hc, Vc_, _ = synth.CONFIGS["C"]
        K, Kinv, sframes = synth.make_stereo_pair(Wc, Hc)
        feats0 = synth.make_features(synth.config_graph("C")["pos"], Wc, Hc)
        ctx.set_camera(Wc, Hc, K, Kinv)
        for i, fr in enumerate(sframes):
            ctx.frame_upload(i, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], synth.PAD, fr["pose"])
        tp = _lib.default_track_params()
        ts = []
        for _ in range(8):
            f = feats0.copy()
            t0 = time.perf_counter()
            n_upd, _ = ctx.update_feature_idepths(tp, 1, 0, f, want_diag=False)
            ts.append(time.perf_counter() - t0)
        t_med = float(np.median(ts[2:]))
        search = {"workload": "config C: 1280x720, %d features, host records in/out" % len(feats0),
                  "ms_per_call": 1e3 * t_med, "features_per_s": len(feats0) / t_med, "updates": int(n_upd),
                  "note": "latency/divergence-bound walk along the epipolar line; no HBM fraction quoted"}

    # ---- e2e: host buffers in, host buffers out, through the public C ABI ----
    e2e = run_e2e(args, local_rank, frames, (W, H, V, iters), barrier)
    e2e_value = shard.aggregate_throughput(e2e["frames"], e2e["seconds"], dev)

    if rank == 0:
        E = int(np.mean([g["edges"].shape[0] for g in frames]))
        peak, peak_src = _peaks()
        alg_bytes = 76.0 * (V + E) * iters * wave  # 76 B per vertex + 76 B per edge per iteration
        achieved = alg_bytes / (pd_ms * 1e-3) / 1e9
        traffic = None
        tp_ = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tp_) and args.mode != 1:  # DRAM bytes of this kernel from the committed ncu capture
            with open(tp_) as f:
                tk = json.load(f)["k_pd_resident"]
            traffic = (tk["dram_bytes_read"] + tk["dram_bytes_write"]) * wave / tk["frames_per_launch"]
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config D: 1920x1080, V=%d, E~%d, %d PD iters + dense projection; "
                                   "%d independent frames per GPU per step, issued as waves of %d (one solver "
                                   "launch per wave = the frame clusters the device keeps resident); the "
                                   "projection of a wave %s" % (V, E, iters, B, wave,
                                                                "runs in stream order" if args.no_async else
                                                                "overlaps the solver of the next wave (second stream)"),
                       "frames_per_step_per_gpu": B, "frames_per_solver_launch": wave,
                       "async_projection": not args.no_async,
                       "pd_mode": ["auto", "stream", "resident"][args.mode],
                       "l2": "per-step working set ~%.0f MB per GPU (PD state 2.4 MB + mesh/bins ~12 MB + 8.3 MB "
                             "dense map per frame) > 126 MB L2: inputs larger than L2" % (B * 23.0),
                       "parallelism": "frames sharded by rank, no data-path collective"},
            "ms_per_pd_iter": pd_ms / iters / wave,
            "single_frame_pd_latency_ms": lat_ms,
            "single_frame_ms_per_pd_iter": lat_ms / iters,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_pd_resident" if args.mode != 1 else "k_dual/primal_stream",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic,
                         "traffic_source": "profiles/r1_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": pd_ms,
                         "note": "algorithmic bytes = 76*(V+E) per PD iteration per frame; the resident kernel "
                                 "keeps the state in shared memory, so achieved is an EFFECTIVE bandwidth "
                                 "(DRAM traffic is far below the algorithmic bytes; see DESIGN.md). The HBM-streaming variant of "
                                 "the same iteration (--mode 1, graphs too large for a cluster) moves 410 MB of DRAM "
                                 "traffic per 128-frame launch at 4.59 TB/s = 0.70 of this peak (ncu, "
                                 "profiles/r1_pd_stream_b128.txt)"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": e2e["h2d"],
                    "d2h_bytes_per_step": e2e["d2h"], "frames_per_step": e2e["per_step"],
                    "pipelines": e2e["pipelines"]},
        }
        if search is not None:
            line["search"] = search
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, local_rank, frames, dims, barrier):
    """Per frame: graph_init + mesh_bind (H2D) -> pd_run -> mesh_project -> dense map + x (D2H).
    `pipelines` host threads each drive their own context/stream so copies overlap compute."""
    from concurrent.futures import ThreadPoolExecutor
    from flame_b200 import _lib
    W, H, V, iters = dims
    P = args.pipelines
    p = _lib.default_pd_params()
    ctxs = [_lib.Context(local_rank, 1) for _ in range(P)]
    outs = [_lib.PinnedArray((H, W)) for _ in range(P)]
    vout = [tuple(np.empty(V, np.float32) for _ in range(3)) for _ in range(P)]
    per_step = len(frames)

    def one(job):
        k, i = job
        g = frames[i]
        ctx, out, (x, w1, w2) = ctxs[k], outs[k], vout[k]
        # the public per-frame call: host arrays in (graph + triangles), host arrays out (dense map
        # into pinned memory, smoothed vertex idepths and plane parameters); synchronises.
        ctx.process_frame(0, g["pos"], g["data_term"], g["edges"], g["tris"], H, W, p, iters, out.array, x, w1, w2)
        return float(x[0]) + float(out.array[H // 2, W // 2])

    pool = ThreadPoolExecutor(P)

    def worker(k, n_steps):
        acc = 0.0
        for s in range(n_steps):
            for i in range(k, per_step, P):
                acc += one((k, i))
        return acc

    def run(n_steps):
        return sum(pool.map(lambda k: worker(k, n_steps), range(P)))

    run(1)  # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    run(args.steps)
    barrier()
    dt = time.perf_counter() - t0
    g = frames[0]
    E, T = g["edges"].shape[0], g["tris"].shape[0]
    # graph_init: pos 2V + data V + weight V floats, 2V+6E ints (+ ~300 of chunk tables); mesh_bind: 12T + T
    h2d = per_step * (4 * (4 * V) + 4 * (2 * V + 6 * E + 300) + 13 * T)
    d2h = per_step * (4 * W * H + 12 * V)
    for c in ctxs:
        c.close()
    return {"seconds": dt, "frames": per_step * args.steps, "h2d": h2d, "d2h": d2h,
            "per_step": per_step, "pipelines": P}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=56, help="frames per GPU per step")
    ap.add_argument("--wave", type=int, default=0,
                    help="frames per solver launch inside a step (0 = what the device keeps resident: pd_wave)")
    ap.add_argument("--no-async", action="store_true", help="projection in stream order (no second stream)")
    ap.add_argument("--pipelines", type=int, default=16, help="host threads / streams for the e2e leg")
    ap.add_argument("--mode", type=int, default=0, help="0 auto, 1 stream, 2 resident")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-search", action="store_true", help="skip the epipolar-search side measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
