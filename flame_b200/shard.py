"""Multi-GPU sharding of the hot path: frames are independent units, so the batch is
sharded BY FRAME across ranks (one process per GPU) with no data-path collective. torch.distributed
is used only for the barrier and for the max-over-ranks timing / sum-over-ranks counters that the
benchmark reports. Works with the `nccl` backend on GPUs and `gloo` on CPU (tests)."""
import os


def rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def frames_for_rank(rank, world, frames_per_rank):
    """Global frame ids [rank*B, (rank+1)*B): weak scaling, every rank owns B distinct frames."""
    if not (0 <= rank < world) or frames_per_rank < 0:
        raise ValueError("bad rank / world / frames_per_rank")
    return list(range(rank * frames_per_rank, (rank + 1) * frames_per_rank))


def frame_owner(frame_id, world, frames_per_rank):
    return (frame_id // frames_per_rank) % world


def max_over_ranks(value, device=None):
    """max over ranks of a python float (elapsed ms); identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(frames_this_rank, elapsed_s_this_rank, device=None):
    """Whole-job frames/s = (sum of frames over ranks) / (max elapsed over ranks)."""
    total = sum_over_ranks(frames_this_rank, device)
    tmax = max_over_ranks(elapsed_s_this_rank, device)
    return total / tmax


def gather_frames(dense_map, values, dst=0):
    """The batched-frame case (BASELINE config E): every rank holds one frame's dense inverse-depth
    map and vertex idepths; rank `dst` receives all of them - NCCL gather over NVLink when the tensors are on
    the GPUs, gloo on CPU tensors. Returns (maps, values) lists of world_size tensors on `dst`, (None, None)
    elsewhere; single-process: the inputs themselves. The vertex counts may differ per rank: values travel
    padded to the largest count, preceded by an all_gather of the counts."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [dense_map], [values]
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([values.numel()], dtype=torch.int64, device=values.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    vmax = max(max(counts), 1)
    padded = torch.zeros(vmax, dtype=values.dtype, device=values.device)
    padded[:values.numel()] = values
    maps = [torch.empty_like(dense_map) for _ in range(world)] if rank == dst else None
    vals = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(dense_map.contiguous(), maps, dst=dst)
    dist.gather(padded, vals, dst=dst)
    if rank != dst:
        return None, None
    return maps, [v[:c] for v, c in zip(vals, counts)]


class PeerGather:
    """Config E without a data-path collective: rank `dst` owns one device buffer [world][rows*cols] (+ [world][V])
    and shares it with the other per-GPU processes of the box through CUDA IPC; every rank redirects its slot's
    projection into its own row (flame_b200_mesh_redirect), so the dense maps land in the consumer's memory
    through peer stores over NVLink while the projection kernel runs. `arrive()` is the completion signal: a
    4-byte NCCL all-reduce issued on the producing streams."""

    def __init__(self, ctx, slot, rows, cols, V, device, dst=0):
        import torch
        import torch.distributed as dist
        from . import _lib
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        self.dst, self.rows, self.cols, self.V = dst, rows, cols, V
        self.map_bytes, self.val_bytes = rows * cols * 4, ((V * 4 + 255) // 256) * 256
        total = self.world * (self.map_bytes + self.val_bytes)
        handle = [None]
        if self.rank == dst:
            self.buf = _lib.DeviceBuffer(device, total)
            handle = [self.buf.export()]
        if self.dist:
            self.dist.broadcast_object_list(handle, src=dst)
        if self.rank != dst:
            self.buf = _lib.DeviceBuffer.open(device, handle[0], total)
        self.flag = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", device))
        ctx.mesh_redirect(slot, self.buf.ptr + self.rank * self.map_bytes,
                          self.buf.ptr + self.world * self.map_bytes + self.rank * self.val_bytes)
        self.ctx, self.slot = ctx, slot

    def arrive(self):
        """Stream-ordered on the caller's current torch stream: completes on `dst` after every rank's projection."""
        if self.dist:
            self.dist.all_reduce(self.flag)

    def results(self):
        """On `dst`: ([world] dense maps, [world] vertex idepths) as views of the gather buffer."""
        if self.rank != self.dst:
            return None, None
        maps = [self.buf.tensor((self.rows, self.cols), r * self.map_bytes) for r in range(self.world)]
        vals = [self.buf.tensor((self.V,), self.world * self.map_bytes + r * self.val_bytes) for r in range(self.world)]
        return maps, vals

    def close(self):
        self.ctx.mesh_redirect(self.slot, None, None)
        self.ctx.sync()
        if self.dist:
            self.dist.barrier()
        self.buf.close()
