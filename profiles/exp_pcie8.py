"""Experiment: the box's aggregate device->host ceiling when N ranks copy at once, and where the pinned
pages live. Launch with torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      profiles/exp_pcie8.py

For every memory policy (default first-touch, preferred = the GPU's NUMA node, interleave over all nodes) each
rank page-locks a 64 x 8.3 MB ring, then (a) all ranks copy 1080p float maps device->host back to back for
~1.5 s, (b) rank 0 and the last rank do the same alone. Rank 0 prints per-rank GB/s and the total.
"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def set_policy(mode, nodes):
    libc = ctypes.CDLL(None, use_errno=True)
    mask = (ctypes.c_ulong * 2)(0, 0)
    for n in nodes:
        mask[n // 64] |= 1 << (n % 64)
    rc = libc.syscall(238, mode, mask, 129)
    return rc == 0


def pages_by_node(addr):
    tag = "%x" % addr
    try:
        with open("/proc/self/numa_maps") as f:
            for ln in f:
                if ln.startswith(tag):
                    return " ".join(t for t in ln.split() if t[0] == "N" and "=" in t)
    except Exception as e:
        return "numa_maps: %s" % e
    return "?"


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import bench
    os.environ["FLAME_B200_MEMPOLICY"] = "default"
    host = bench._pin_to_gpu_node(local)
    nodes = sorted(int(n[4:]) for n in os.listdir("/sys/devices/system/node") if n.startswith("node"))
    mems = "?"
    with open("/proc/self/status") as f:
        for ln in f:
            if ln.startswith("Mems_allowed_list"):
                mems = ln.split()[1]
    aff = sorted(os.sched_getaffinity(0))
    me = "rank %d: gpu node %s, cpus %d (%d..%d) cpu node %s, mems allowed %s, nodes %s" % (
        rank, host.get("gpu_numa_node"), len(aff), aff[0], aff[-1], host.get("numa_node"), mems, nodes)
    gathered = [None] * world
    dist.all_gather_object(gathered, me)
    if rank == 0:
        print("\n".join(gathered), flush=True)

    nbytes = 1920 * 1080 * 4
    R = 64
    src = torch.empty(4, nbytes, dtype=torch.uint8, device=dev).fill_(rank + 1)
    gnode = host.get("gpu_numa_node")
    policies = [("default", 0, [])]
    if gnode is not None and gnode >= 0:
        policies.append(("gpu-node", 1, [gnode]))
    if len(nodes) > 1:
        policies.append(("interleave", 3, nodes))
    st = torch.cuda.Stream(dev)
    for name, mode, ns in policies:
        ok = set_policy(mode, ns)
        ring = torch.empty(R, nbytes, dtype=torch.uint8).pin_memory()
        ring.fill_(0)
        where = pages_by_node(ring.data_ptr())
        set_policy(0, [])

        def burst(seconds):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 0
            with torch.cuda.stream(st):
                a.record(st)
                t0 = time.perf_counter()
                while time.perf_counter() - t0 < seconds:
                    for _ in range(16):
                        ring[n % R].copy_(src[n % 4], non_blocking=True)
                        n += 1
                    st.synchronize()
                b.record(st)
            st.synchronize()
            return n * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9

        burst(0.2)
        dist.barrier()
        together = burst(1.5)
        dist.barrier()
        alone = None
        for r in (0, world - 1):
            if rank == r:
                alone = burst(1.0)
            dist.barrier()
        rows = [None] * world
        dist.all_gather_object(rows, (rank, ok, where, together, alone))
        if rank == 0:
            print("policy %s:" % name)
            for r, ok_, where_, tg, al in rows:
                print("  rank %d set=%s pages[%s]  together %.1f GB/s%s" % (
                    r, ok_, where_, tg, "" if al is None else "  alone %.1f GB/s" % al))
            print("  total together: %.1f GB/s" % sum(x[3] for x in rows), flush=True)
        del ring
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
