"""Worker of tests/test_multi_gpu.py, one process per GPU (torchrun): rank r solves and projects frame r on
cuda:r through the C ABI, the dense maps and vertex idepths are gathered to rank 0 with NCCL, and rank 0
checks EVERY rank's result bit for bit against the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from flame_b200 import _lib, shard, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = "B"
    W, H, V, iters = synth.CONFIGS[cfg]
    g = synth.config_graph(cfg, frame=rank)
    p = _lib.default_pd_params()
    with _lib.Context(local, 1) as ctx:
        ctx.graph_init(0, g["pos"], g["data_term"], g["edges"])
        ctx.mesh_bind(0, g["tris"], H, W)
        ctx.pd_run(p, iters)
        ctx.mesh_project(0, 1, 1.0)
        stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        m, v = ctx.device_tensors(0)
        with torch.cuda.stream(stream):
            maps, vals = shard.gather_frames(m, v, dst=0)
        ctx.sync()
        torch.cuda.synchronize(dev)
        # second exchange path: the projection stores straight into rank 0's buffer (peer memory, CUDA IPC)
        V_all = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(V_all, torch.tensor([g["pos"].shape[0]], dtype=torch.int64, device=dev))
        Vmax = max(int(x.item()) for x in V_all)
        pg = shard.PeerGather(ctx, 0, H, W, Vmax, local)
        ctx.mesh_project(0, 1, 1.0)
        with torch.cuda.stream(stream):
            pg.arrive()
        ctx.sync()
        torch.cuda.synchronize(dev)
        pmaps, pvals = pg.results()
        ok = True
        if rank == 0:
            from oracle import pyoracle
            for r in range(world):
                vr = int(V_all[r].item())
                if not (torch.equal(pmaps[r].view(torch.int32), maps[r].view(torch.int32)) and
                        torch.equal(pvals[r][:vr].view(torch.int32), vals[r].view(torch.int32))):
                    print("rank %d: peer-memory exchange DIFFERS from the NCCL gather" % r)
                    ok = False
                gr = synth.config_graph(cfg, frame=r)
                alpha, beta = pyoracle.edge_weights(gr["edges"], gr["pos"])
                s = synth.init_state(gr, alpha, beta)
                pyoracle.pd_run(s, pyoracle.pd_default_params(), iters)
                img = pyoracle.interpolate_mesh(gr["tris"], gr["pos"], s["x"], H, W)
                got_img, got_x = maps[r].cpu().numpy(), vals[r].cpu().numpy()
                same_x = np.array_equal(got_x.view(np.uint32), s["x"].view(np.uint32))
                nan_g, nan_w = np.isnan(got_img), np.isnan(img)
                same_img = np.array_equal(nan_g, nan_w) and np.array_equal(
                    got_img[~nan_g].view(np.uint32), img[~nan_w].view(np.uint32))
                print("rank %d: x %s, dense map %s" % (r, "ok" if same_x else "DIFFERS", "ok" if same_img else "DIFFERS"))
                ok = ok and same_x and same_img
        dist.barrier()
        pg.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK" if ok else "MGPU_FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
