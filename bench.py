#!/usr/bin/env python
"""bench.py — headline benchmark of the FLaME variational-depth hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config B|C|D]

Workload (default: BASELINE.json configs[3], the one `metric` is quoted on): 1920x1080 frames, 10 000
Delaunay vertices each, 200 primal-dual iterations + dense inverse-depth projection. `--config B` / `C`
run configs[1] / configs[2] (640x480 / 2k vertices / 100 iterations; 1280x720 / 5k vertices / epipolar
search + 100 iterations) the same way.
A "step" is one pass of the hot path over one batch of independent synthetic frames per GPU:
    [epipolar search (C)] -> graph_reset (fresh-vertex rule) -> n x step() -> interpolateMesh
`value`  = frames/s with inputs resident in HBM (CUDA events on the library's stream, max over ranks).
`e2e`    = the same metric through the public C ABI with HOST buffers: per frame, upload of the
           graph + triangles (H2D), PD, projection, download of the dense map + vertex idepths (D2H),
           as a pipeline of frames driven through flame_b200_process_frame_submit / _wait.
`roofline` is for the dominant kernel (the resident PD solver; `--mode 1` = the HBM-streaming solver).
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
_METRICS = {
    "D": "frames/sec at 1080p (10k vertices, 200 PD iters + dense projection)",
    "C": "frames/sec at 720p (5k vertices, epipolar search + 100 PD iters + dense projection)",
    "B": "frames/sec at VGA (2k vertices, 100 PD iters + dense projection)",
    "E": "frames/sec at 1080p, one frame per GPU (10k vertices, 200 PD iters + dense projection), results gathered "
         "to rank 0 over NVLink",
}
_WORK = {
    "D": "config D: 1920x1080, V=10000, 200 PD iters + dense projection",
    "C": "config C: 1280x720, V=5000, epipolar search + 100 PD iters + dense projection",
    "B": "config B: 640x480, V=2000, 100 PD iters + dense projection",
    "E": "config E: batched 1920x1080 frames, one per GPU, V=10000, 200 PD iters + dense projection, dense maps + "
         "vertex idepths gathered to rank 0 with NCCL",
}
N_DISTINCT = 30  # distinct synthetic frames; the slots of a batch cycle through them


def _metric():
    return _METRICS[CFG]


def _config():
    """The `config` object is the same in both arms (the driver compares them)."""
    return {"workload": _WORK[CFG]}


def _usable_cpus():
    """CPUs this process may really use: its affinity mask, capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0))
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, n)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _frames(n, first=0):
    """n frame descriptions; at most N_DISTINCT distinct ones (generated once), cycled."""
    from flame_b200 import synth
    cfg = "D" if CFG == "E" else CFG  # config E = config D's frames, one per GPU
    W, H, V, iters = synth.CONFIGS[cfg]
    distinct = [synth.config_graph(cfg, frame=first + f) for f in range(min(n, N_DISTINCT))]
    return [distinct[i % len(distinct)] for i in range(n)], (W, H, V, iters)


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
                time.sleep(0.01)
        except Exception as e:  # pragma: no cover
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


_SEARCH = {}


def _search_inputs():
    """Config C's stereo pair + features (the epipolar search that seeds the frame), built once."""
    if not _SEARCH:
        from flame_b200 import synth
        Wc, Hc, _, _ = synth.CONFIGS["C"]
        K, Kinv, sframes = synth.make_stereo_pair(Wc, Hc)
        feats0 = synth.make_features(synth.config_graph("C")["pos"], Wc, Hc)
        _SEARCH.update(W=Wc, H=Hc, K=K, Kinv=Kinv, frames=sframes, feats=feats0)
    return _SEARCH


def _cpu_frame(pyoracle, synth, g, iters):
    """One frame of the hot path on one CPU thread (the reference runs step() on exactly one thread,
    flame.cc:100-111): [epipolar search (config C)], edge weights, n x step, interpolateMesh."""
    if CFG == "C":
        si = _search_inputs()
        fn = pyoracle.ref_update_feature_idepths if pyoracle.have_ref() else pyoracle.update_feature_idepths
        fn(pyoracle.track_default_params(), si["K"], si["Kinv"], si["W"], si["H"], synth.PAD, si["frames"], 1, 0,
           si["feats"].copy())
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
           "sample": "%d frames of %s, %s, 1 thread (the reference runs step() on one thread), %.1f s"
                     % (n, _WORK[CFG], _CPU_DESC[_cpu_kind(pyoracle)], dt)}
    try:  # the epipolar search alone on one CPU thread (OpenMP is inert in the reference build)
        si = _search_inputs()
        fn = pyoracle.ref_update_feature_idepths if pyoracle.have_ref() else pyoracle.update_feature_idepths
        tp = pyoracle.track_default_params()
        ts = []
        for _ in range(3):
            f = si["feats"].copy()
            t0 = time.perf_counter()
            fn(tp, si["K"], si["Kinv"], si["W"], si["H"], synth.PAD, si["frames"], 1, 0, f)
            ts.append(time.perf_counter() - t0)
        out["search"] = {"ms_per_call": 1e3 * min(ts), "features_per_s": len(si["feats"]) / min(ts), "cores": 1,
                         "workload": "config C: %dx%d, %d features" % (si["W"], si["H"], len(si["feats"]))}
    except Exception as e:  # the headline baseline above does not depend on it
        out["search"] = {"unavailable": repr(e)}
    return out


def _ref_worker(conn, cfg):
    """One worker PROCESS = one Flame instance on one core: runs whole frames until told to stop."""
    global CFG
    CFG = cfg
    from flame_b200 import synth
    from oracle import pyoracle
    frames, (_, _, _, iters) = _frames(1)
    g = frames[0]
    if CFG == "C":
        _search_inputs()
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
    parallel, one single-threaded worker process per USABLE core (each Flame instance runs step() on one
    thread, flame.cc:100-111). A step = one frame per core."""
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import pyoracle
    pyoracle.build()
    cores = _usable_cpus()
    ctx = mp.get_context("fork")
    workers = []
    for _ in range(cores):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(b, CFG), daemon=True)
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
        "impl": "reference", "metric": _metric(), "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": _config(),
        "details": {"frames_per_step": per_step, "os_cpu_count": os.cpu_count(),
                    "affinity_cpus": len(os.sched_getaffinity(0))},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": _cpu_kind(pyoracle),
                         "per_core": val / cores,
                         "sample": "%d frames per step (one per usable core = affinity mask capped by the cgroup CPU "
                                   "quota; one single-threaded worker process per core, as each Flame instance runs "
                                   "step() on one thread), %s" % (per_step, _CPU_DESC[_cpu_kind(pyoracle)])},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def _prefer_node(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) for this process (x86-64 syscall 238); returns what happened."""
    import ctypes
    libc = ctypes.CDLL(None, use_errno=True)
    mask = (ctypes.c_ulong * 2)(0, 0)
    mask[node // 64] = 1 << (node % 64)
    rc = libc.syscall(238, 1, mask, 129)
    return "preferred node %d" % node if rc == 0 else "refused (errno %d)" % ctypes.get_errno()


def _pin_to_gpu_node(index):
    """Run this rank on the CPUs of its GPU's NUMA node (nvml's ideal CPU affinity of the device): host
    threads and the pinned buffers they allocate (first touch) then sit next to the GPU's PCIe root.
    Returns (cpus in the mask, numa node of the first one) for the bench line."""
    info = {"cpus": len(os.sched_getaffinity(0)), "numa_node": None, "pinned": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        # the GPU's own NUMA node (sysfs): the page-locked result ring is allocated there even when the
        # job's CPU mask does not reach that socket (set_mempolicy(MPOL_PREFERRED); a no-op where the
        # container's memory nodes exclude it)
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            with open("/sys/bus/pci/devices/%s/numa_node" % bus[-12:].lower()) as f:
                info["gpu_numa_node"] = int(f.read())
            if info["gpu_numa_node"] >= 0 and os.environ.get("FLAME_B200_MEMPOLICY", "gpu") == "gpu":
                info["mempolicy"] = _prefer_node(info["gpu_numa_node"])
        except Exception:
            pass
        n = os.cpu_count()
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = [i for i in range(n) if (words[i // 64] >> (i % 64)) & 1 and i in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = len(cpus)
            info["pinned"] = True
            for node in sorted(os.listdir("/sys/devices/system/node")):
                if node.startswith("node") and os.path.exists("/sys/devices/system/node/%s/cpu%d" % (node, cpus[0])):
                    info["numa_node"] = int(node[4:])
                    break
    except Exception:
        pass
    return info


def run_ours(args, rank, world, local_rank):
    import torch
    from flame_b200 import _lib, synth

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa = _pin_to_gpu_node(local_rank)

    from flame_b200 import shard
    W, H, V, iters = synth.CONFIGS[CFG]
    p = _lib.default_pd_params()
    if args.e2e_only:
        def barrier0():
            torch.cuda.synchronize(dev)
            if dist is not None:
                dist.barrier()
        fr, _ = _frames(N_DISTINCT)
        e = run_e2e(args, local_rank, fr, (W, H, V, iters), barrier0)
        print(json.dumps({"rank": rank, "e2e_frames_per_s": e["frames"] / e["seconds"], "threads": e["threads"],
                          "in_flight": e["in_flight"], "breakdown": e["breakdown"], "host": numa}))
        return
    # ---- how many frames the device solves concurrently (one thread-block cluster per frame) ----
    pref = _lib.PREF_LATENCY if args.layout == "wide" else _lib.PREF_THROUGHPUT
    probe, _ = _frames(1)
    with _lib.Context(local_rank, 1) as c0:
        c0.set_pd_preference(pref)
        c0.graph_init(0, probe[0]["pos"], probe[0]["data_term"], probe[0]["edges"])
        wave = c0.pd_wave(0) if args.mode != 1 else 0
    if args.wave > 0:
        wave = args.wave
    B = args.batch if args.batch > 0 else (40 * wave if wave > 0 else 128)
    if wave <= 0:
        wave = B
    wave = min(wave, B)
    ids = shard.frames_for_rank(rank, world, B)  # frames shard by rank: no data-path collective
    frames, _ = _frames(B, first=ids[0] % 1000)
    ctx = _lib.Context(local_rank, B)
    ctx.set_pd_preference(pref)
    for i, g in enumerate(frames):
        ctx.graph_init(i, g["pos"], g["data_term"], g["edges"])
        ctx.mesh_bind(i, g["tris"], H, W)
    ctx.sync()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    # A frame is one cluster; the device keeps `wave` of them resident. A step walks its B frames wave by
    # wave; the projection of a wave runs on the context's second stream and overlaps the solver of the next.
    ctx.set_async_projection(not args.no_async)
    waves = [(s0, min(wave, B - s0)) for s0 in range(0, B, wave)]
    search_dev = None
    if CFG == "C":  # the epipolar search is part of the frame: device-resident features, one launch per frame
        si = _search_inputs()
        ctx.set_camera(si["W"], si["H"], si["K"], si["Kinv"])
        for i, fr in enumerate(si["frames"]):
            ctx.frame_upload(i, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], synth.PAD, fr["pose"])
        tp = _lib.default_track_params()
        ctx.features_upload(si["feats"])
        search_dev = (tp, len(si["feats"]))

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(ev=None):
        for w, (s0, n) in enumerate(waves):
            if search_dev is not None:
                for _ in range(n):  # one search per frame of the wave (5 000 device-resident features each)
                    ctx.features_track(search_dev[0], 1, 0)
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
    ctx.set_async_projection(False)

    # ---- single-frame latency: the latency layout (frame spread over up to 16 SMs), B = 1 ----
    lat_ms = None
    if args.mode != 1:
        with _lib.Context(local_rank, 1) as cl:
            cl.set_pd_preference(_lib.PREF_LATENCY)
            cl.graph_init(0, frames[0]["pos"], frames[0]["data_term"], frames[0]["edges"])
            ls = torch.cuda.ExternalStream(cl.stream, device=dev)
            lat = []
            for _ in range(6):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                cl.graph_reset(0, 1)
                a.record(ls)
                cl.pd_run(p, iters, 0, 1, args.mode)
                b.record(ls)
                cl.sync()
                lat.append(a.elapsed_time(b))
            lat_ms = float(np.median(lat[1:]))

    # ---- epipolar search (config C: 1280x720, 5000 features), reported as features/s ----
    search = None
    if rank == 0 and not args.no_search:
        si = _search_inputs()
        ctx.set_camera(si["W"], si["H"], si["K"], si["Kinv"])
        for i, fr in enumerate(si["frames"]):
            ctx.frame_upload(i, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], synth.PAD, fr["pose"])
        tp = _lib.default_track_params()
        ts = []
        for _ in range(8):
            f = si["feats"].copy()
            t0 = time.perf_counter()
            n_upd, _ = ctx.update_feature_idepths(tp, 1, 0, f, want_diag=False)
            ts.append(time.perf_counter() - t0)
        t_med = float(np.median(ts[2:]))
        search = {"workload": "config C: 1280x720, %d features" % len(si["feats"]),
                  "host_call_ms": 1e3 * t_med, "host_call_features_per_s": len(si["feats"]) / t_med,
                  "updates": int(n_upd)}
        # device-timed: features resident, CUDA events around 20 launches
        ctx.features_upload(si["feats"])
        ctx.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        for _ in range(3):
            ctx.features_track(tp, 1, 0)
        a.record(stream)
        for _ in range(reps):
            ctx.features_track(tp, 1, 0)
        b.record(stream)
        ctx.sync()
        k_ms = a.elapsed_time(b) / reps
        search.update({"kernel_ms": k_ms, "features_per_s": len(si["feats"]) / (k_ms * 1e-3),
                       "note": "device-timed (CUDA events, features resident in HBM: geometry-table + tracking kernel); "
                               "one warp per feature, patches staged in shared memory; latency-bound, no HBM fraction quoted"})

    # ---- e2e: host buffers in, host buffers out, through the public C ABI ----
    ctx.close()
    e2e = run_e2e(args, local_rank, frames, (W, H, V, iters), barrier)
    e2e_value = shard.aggregate_throughput(e2e["frames"], e2e["seconds"], dev)
    e2e_tri = None
    if not args.no_tri:
        # the same pipeline with the Delaunay triangulation of every frame done by the host threads
        # (flame_b200_delaunay) - the reference pays Triangle per frame too (utils/delaunay.cc:31-77)
        e2e_tri = run_e2e(args, local_rank, frames, (W, H, V, iters), barrier, triangulate=True)
        e2e_tri["value"] = shard.aggregate_throughput(e2e_tri["frames"], e2e_tri["seconds"], dev)

    ceiling = _d2h_ceiling(torch, dev, dist, barrier, 4 * W * H)

    if rank == 0:
        E = int(np.mean([g["edges"].shape[0] for g in frames[:N_DISTINCT]]))
        peak, peak_src = _peaks()
        # 76 B per vertex + 76 B per edge per iteration
        if args.mode == 1:
            kernel, alg_bytes, launch_ms = "k_pd_stream", 76.0 * (V + E) * wave, pd_ms / iters
            tkey = "k_pd_stream"
        else:
            kernel, alg_bytes, launch_ms = "k_pd_resident", 76.0 * (V + E) * iters * wave, pd_ms
            tkey = "k_pd_resident"
        achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
        traffic, tsrc = None, None
        tp_ = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tp_) and CFG == "D":  # DRAM bytes of this kernel from the committed ncu capture
            with open(tp_) as f:
                tk = json.load(f).get(tkey)
            if tk:
                traffic = (tk["dram_bytes_read"] + tk["dram_bytes_write"]) * wave / tk["frames_per_launch"]
                tsrc = "profiles/r2_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, %s)" % tk["capture"]
        line = {
            "metric": _metric(), "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _config(),
            "details": {"E": E, "frames_per_step_per_gpu": B, "frames_per_solver_launch": wave,
                        "distinct_synthetic_frames": min(B, N_DISTINCT),
                        "async_projection": not args.no_async,
                        "pd_mode": ["auto", "stream", "resident"][args.mode], "cluster_layout": args.layout,
                        "issue": "a step walks its frames in waves of %d (one solver launch per wave = the frame "
                                 "clusters the device keeps resident); the projection of a wave %s" % (
                                     wave, "runs in stream order" if args.no_async else
                                     "overlaps the solver of the next wave (second stream)"),
                        "l2": "per-step working set ~%.1f GB per GPU (PD state + mesh + dense map per frame) > 126 MB L2: "
                              "inputs larger than L2" % (B * (4.0 * W * H * 2 + 300.0 * V) / 1e9),
                        "parallelism": "frames sharded by rank, no data-path collective",
                        "host": numa},
            "ms_per_pd_iter": pd_ms / iters / wave,
            "single_frame_pd_latency_ms": lat_ms,
            "single_frame_ms_per_pd_iter": (lat_ms / iters) if lat_ms else None,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                         "traffic_source": tsrc, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": launch_ms,
                         "note": ("algorithmic bytes = 76*(V+E) per PD iteration per frame; one launch = one iteration of "
                                  "%d frames with the state in HBM" % wave) if args.mode == 1 else
                                 ("algorithmic bytes = 76*(V+E) per PD iteration per frame; the resident kernel keeps the "
                                  "state in shared memory for all %d iterations, so achieved is an EFFECTIVE bandwidth "
                                  "(DRAM traffic is far below the algorithmic bytes; see DESIGN.md). `--mode 1 --batch 128` "
                                  "prints the same line for the HBM-streaming solver, whose HBM fraction ncu evidences" % iters)},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": e2e["h2d"],
                    "d2h_bytes_per_step": e2e["d2h"], "frames_per_step": e2e["per_step"],
                    "host_threads": e2e["threads"], "frames_in_flight": e2e["in_flight"],
                    "host_breakdown": e2e["breakdown"],
                    "pcie_d2h_ceiling": dict(ceiling, frames_per_s=ceiling["GBps_total"] * 1e9 / (e2e["d2h"] / e2e["per_step"]),
                                             note="all ranks copying dense maps device->pinned host back to back, nothing "
                                                  "else running: what the box's PCIe fabric gives the e2e leg (every "
                                                  "frame returns 4*W*H + 12*V bytes)")},
        }
        if e2e_tri is not None:
            line["e2e_incl_triangulation"] = {
                "value": e2e_tri["value"], "unit": "frames/s", "host_threads": e2e_tri["threads"],
                "delaunay_ms_per_frame": 1e-3 * e2e_tri["breakdown"]["per_frame_us"]["delaunay"],
                "note": "every frame's Delaunay triangulation (10k vertices) computed inside "
                        "flame_b200_process_frame_submit (no topology passed: flame_b200_delaunay on the submitting host "
                        "thread) while earlier frames run on the GPU; host-core bound"}
        if search is not None:
            line["search"] = search
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_config_e(args, rank, world, local_rank):
    """BASELINE config E: the batched-frame case. Every GPU holds ONE 1080p / 10k-vertex frame in
    the latency layout (the frame spread over a 16-CTA cluster); a step = fresh-vertex reset -> 200 x step() ->
    interpolateMesh on every GPU, then the dense maps and vertex idepths are gathered to rank 0 with NCCL over
    NVLink. `value` = frames/s of the whole job, next to the per-step latency and the gather's bandwidth."""
    import torch
    import torch.distributed as dist
    from flame_b200 import _lib, shard, synth
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    numa = _pin_to_gpu_node(local_rank)
    W, H, V, iters = synth.CONFIGS["D"]
    g = synth.config_graph("D", frame=rank)
    p = _lib.default_pd_params()
    ctx = _lib.Context(local_rank, 1)
    ctx.set_pd_preference(_lib.PREF_LATENCY)
    ctx.graph_init(0, g["pos"], g["data_term"], g["edges"])
    ctx.mesh_bind(0, g["tris"], H, W)
    ctx.sync()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    m, v = ctx.device_tensors(0)
    # exchange step: "peer" = the projection kernel stores its tiles directly into rank 0's gather buffer (CUDA IPC
    # peer memory over NVLink) and a 4-byte all-reduce signals completion; "nccl" = ncclGather of the finished maps
    pg = shard.PeerGather(ctx, 0, H, W, V, local_rank) if args.exchange == "peer" else None

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(ev=None):
        ctx.graph_reset(0, 1)
        ctx.pd_run(p, iters, 0, 1, 0)
        if ev is not None and pg is not None:
            ev[0].record(stream)  # peer mode: the exchange IS the projection
        ctx.mesh_project(0, 1, 1.0)
        if ev is not None and pg is None:
            ev[0].record(stream)
        with torch.cuda.stream(stream):
            if pg is not None:
                pg.arrive()
                out = pg.results()
            else:
                out = shard.gather_frames(m, v, dst=0)
        if ev is not None:
            ev[1].record(stream)
        return out

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.kernel_launches()
    e0.record(stream)
    for k in range(args.steps):
        maps, vals = step(evs[k])
    e1.record(stream)
    barrier()
    launches = _lib.kernel_launches() - launches0
    ms_total = shard.max_over_ranks(e0.elapsed_time(e1), dev)
    gather_ms = shard.max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in evs])), dev)
    clocks = sampler.stop()
    checksum = None
    if rank == 0:
        checksum = float(sum(torch.nan_to_num(x).double().sum().item() for x in maps))
    if rank == 0:
        E = g["edges"].shape[0]
        peak, peak_src = _peaks()
        gathered = (world - 1) * (4.0 * W * H + 4.0 * V)  # bytes arriving at rank 0 per step
        solve_ms = ms_total / args.steps - gather_ms
        line = {
            "metric": _metric(), "value": world * args.steps / (ms_total * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(),
            "details": {"E": E, "frames_per_step_per_gpu": 1, "layout": "latency (16-CTA cluster per frame)",
                        "exchange": "projection stores into rank 0's buffer through peer memory (NVLink) + 4-byte "
                                    "all-reduce; gather_ms = projection + signal" if pg is not None else
                                    "ncclGather of the dense maps and vertex idepths; gather_ms = the collective",
                        "latency_ms_per_batch": ms_total / args.steps, "gather_ms": gather_ms,
                        "gather_GBps_into_rank0": gathered / (gather_ms * 1e-3) / 1e9 if world > 1 else None,
                        "gathered_bytes_per_step": gathered, "checksum": checksum, "host": numa,
                        "parallelism": "one frame per GPU; NCCL gather of the results to rank 0"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_pd_resident", "achieved": 76.0 * (V + E) * iters / (solve_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": 76.0 * (V + E) * iters / (solve_ms * 1e-3) / 1e9 / peak,
                         "peak_source": peak_src, "traffic": None,
                         "note": "single frame per GPU: effective bandwidth of solve + projection (step minus gather)"},
            "e2e": {"value": world * args.steps / (ms_total * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0,
                    "note": "config E keeps inputs and outputs on the GPUs (consumer = rank 0's GPU over NVLink)"},
        }
        print(json.dumps(line))
    if pg is not None:
        pg.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _d2h_ceiling(torch, dev, dist, barrier, nbytes, seconds=0.6):
    """Raw device->host rate of this box with every rank copying at once (GB/s: total, slowest and fastest rank)."""
    src = torch.empty(4, nbytes, dtype=torch.uint8, device=dev)
    ring = torch.empty(16, nbytes, dtype=torch.uint8).pin_memory()
    st = torch.cuda.Stream(dev)

    def burst(sec):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 0
        with torch.cuda.stream(st):
            a.record(st)
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < sec:
                for _ in range(8):
                    ring[n % 16].copy_(src[n % 4], non_blocking=True)
                    n += 1
                st.synchronize()
            b.record(st)
        st.synchronize()
        return n * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9

    burst(0.1)
    barrier()
    mine = burst(seconds)
    barrier()
    if dist is None:
        return {"GBps_total": mine, "GBps_slowest_rank": mine, "GBps_fastest_rank": mine}
    t = torch.tensor([mine, -mine, mine], dtype=torch.float64, device=dev)
    tot = t.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return {"GBps_total": float(tot[0]), "GBps_slowest_rank": -float(mx[1]), "GBps_fastest_rank": float(mx[2])}


def run_e2e(args, local_rank, frames, dims, barrier, triangulate=False):
    """Per frame: graph + triangles up (H2D) -> PD -> projection -> dense map + x, w1, w2 down (D2H), through
    flame_b200_process_frame_submit / _wait. `pipelines` host threads, each driving ONE context with a ring of
    `ring` slots: while the device works on the frames in flight, the thread prepares and submits the next."""
    from concurrent.futures import ThreadPoolExecutor
    from flame_b200 import _lib
    W, H, V, iters = dims
    P, R = args.pipelines, args.ring
    if triangulate:  # the host also triangulates every frame (flame_b200_delaunay): many threads, short rings
        P, R = args.tri_threads, 6
    p = _lib.default_pd_params()
    ctxs = [_lib.Context(local_rank, R) for _ in range(P)]
    outs = [[_lib.PinnedArray((H, W)) for _ in range(R)] for _ in range(P)]
    vout = [tuple(np.empty(V, np.float32) for _ in range(3)) for _ in range(P)]
    per_step = args.e2e_batch if not triangulate else max(P * 8, 96)
    search = CFG == "C"
    if search:
        from flame_b200 import synth
        si = _search_inputs()
        tp = _lib.default_track_params()
        for c in ctxs:
            c.set_camera(si["W"], si["H"], si["K"], si["Kinv"])
            for i, fr in enumerate(si["frames"]):
                c.frame_upload(i, fr["img_pad"], fr["gradx_pad"], fr["grady_pad"], synth.PAD, fr["pose"])

    def worker(k, n_steps):
        ctx, (x, w1, w2) = ctxs[k], vout[k]
        busy = [False] * R
        acc, j = 0.0, 0
        for s in range(n_steps):
            for i in range(k, per_step, P):
                slot = j % R
                if busy[slot]:
                    ctx.process_frame_wait(slot, x, w1, w2)
                    acc += float(x[0]) + float(outs[k][slot].array[H // 2, W // 2])
                g = frames[i % len(frames)]
                tris_i, edges_i = g["tris"], g["edges"]
                if triangulate:  # no topology passed: process_frame_submit triangulates this frame's vertices itself,
                    tris_i = edges_i = None  # on this host thread, while the frames submitted before run on the GPU
                if search:  # host feature records in and out (5 000 x 40 B each way), synchronous
                    f = si["feats"].copy()
                    ctx.update_feature_idepths(tp, 1, 0, f, want_diag=False)
                ctx.process_frame_submit(slot, g["pos"], g["data_term"], edges_i, tris_i, H, W, p,
                                         int(os.environ.get("E2E_ITERS", iters)),
                                         None if os.environ.get("E2E_NOMAP") else outs[k][slot].array)
                busy[slot] = True
                j += 1
        for slot in range(R):
            if busy[slot]:
                ctx.process_frame_wait(slot, x, w1, w2)
                acc += float(x[0]) + float(outs[k][slot].array[H // 2, W // 2])
        return acc

    pool = ThreadPoolExecutor(P)

    def run(n_steps):
        return sum(pool.map(lambda k: worker(k, n_steps), range(P)))

    run(1)  # warm-up (allocations)
    for c in ctxs:
        c.stats(reset=True)
    barrier()
    t0 = time.perf_counter()
    run(args.steps)
    barrier()
    dt = time.perf_counter() - t0
    st = [c.stats() for c in ctxs]
    nfr = max(1, sum(x["frames"] for x in st))
    tot = lambda key: sum(x[key] for x in st)
    g = frames[0]
    E, T = g["edges"].shape[0], g["tris"].shape[0]
    # graph: pos 2V + data V + weight V floats, vertex order V + edges 2E ints (+ chunk tables); mesh: 12T
    h2d = per_step * (4 * (4 * V) + 4 * (V + 2 * E + 300) + 12 * T + (5000 * 40 if search else 0))
    d2h = per_step * (4 * W * H + 12 * V + (5000 * 40 if search else 0))
    breakdown = {
        "per_frame_us": {"vertex_order_and_plan": 1e6 * tot("prep_seconds") / nfr,
                         "staging_and_upload_calls": 1e6 * (tot("stage_seconds") - tot("prep_seconds")) / nfr,
                         "launches_and_download_calls": 1e6 * tot("enqueue_seconds") / nfr,
                         "blocked_in_wait": 1e6 * tot("wait_seconds") / nfr},
        "host_thread_busy_frac": (tot("stage_seconds") + tot("enqueue_seconds")) / (P * dt),
        "h2d_GBps": h2d * args.steps / dt / 1e9, "d2h_GBps": d2h * args.steps / dt / 1e9,
    }
    for c in ctxs:
        c.close()
    for row in outs:
        for o in row:
            o.free()
    if triangulate:
        breakdown["per_frame_us"]["delaunay"] = 1e6 * tot("triangulation_seconds") / nfr
    return {"seconds": dt, "frames": per_step * args.steps, "h2d": h2d, "d2h": d2h,
            "per_step": per_step, "threads": P, "in_flight": P * R, "breakdown": breakdown}


def main():
    global CFG
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="D", choices=["B", "C", "D", "E"])
    ap.add_argument("--batch", type=int, default=0,
                    help="frames per GPU per step (0 = 40 waves of what the device keeps resident)")
    ap.add_argument("--wave", type=int, default=0,
                    help="frames per solver launch inside a step (0 = what the device keeps resident: pd_wave)")
    ap.add_argument("--no-async", action="store_true", help="projection in stream order (no second stream)")
    ap.add_argument("--layout", default="compact", choices=["wide", "compact"],
                    help="device-resident leg: the cluster size that keeps most SMs busy (9 CTAs per frame, 15 frames "
                         "per wave on 135 SMs; the low-priority projection stream fills the rest and the gaps) or "
                         "frame clusters of 16 CTAs (7 frames per wave, lower latency per frame)")
    ap.add_argument("--pipelines", type=int, default=0,
                    help="host threads (one context each) of the e2e leg (0 = 3; 4 at three or more ranks if every rank "
                         "has at least 6 usable CPUs: the per-frame host work - vertex order, launch plan, staging - slows "
                         "down with the box's DMA load, and at 4 ranks three threads per rank were measured host-bound, "
                         "profiles/r2_bench_n4_final.json; at one rank a fourth thread only adds contention)")
    ap.add_argument("--ring", type=int, default=30, help="frame slots per e2e host thread (two 15-frame waves)")
    ap.add_argument("--e2e-batch", type=int, default=480, help="frames per GPU per step of the e2e leg")
    ap.add_argument("--mode", type=int, default=0, help="0 auto, 1 stream, 2 resident")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-search", action="store_true", help="skip the epipolar-search side measurement")
    ap.add_argument("--no-tri", action="store_true", help="skip the e2e leg that includes the host triangulation")
    ap.add_argument("--tri-threads", type=int, default=0,
                    help="host threads of the e2e leg that triangulates (0 = usable cpus / ranks on the box - 2, <= 12)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="config E: how the results reach rank 0 (peer stores from the projection kernel / ncclGather)")
    ap.add_argument("--e2e-only", action="store_true", help="experiments: only the e2e leg (prints its dict)")
    args = ap.parse_args()
    CFG = args.config
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.pipelines <= 0:
        args.pipelines = 4 if (world >= 3 and _usable_cpus() // world >= 6) else 3
    if args.tri_threads <= 0:
        args.tri_threads = max(2, min(12, _usable_cpus() // max(1, world) - 2))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if CFG == "E":
        run_config_e(args, rank, world, local_rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
