"""Multi-GPU plumbing of the frame-parallel path: one process per GPU, no data-path collective.

The hot path shards by independent sensor streams / fresh-state frames: rank r of W owns the
contiguous block of global frame (stream) indices [r*F, (r+1)*F) when every rank processes F frames (weak scaling),
or the balanced contiguous split of a fixed total (strong scaling). torch.distributed is used only for the
rendezvous, the barriers around the timed region and the max-over-ranks of the elapsed time (NCCL on GPUs, gloo in
the CPU tests); results stay on the rank that produced them.
"""
import os

import torch


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def weak_shard(frames_per_rank: int, rank: int, world: int):
    """Global frame indices of `rank` when every rank processes `frames_per_rank` frames."""
    assert 0 <= rank < world and frames_per_rank >= 0
    return range(rank * frames_per_rank, (rank + 1) * frames_per_rank)


def strong_shard(total_frames: int, rank: int, world: int):
    """Balanced contiguous split of a fixed number of frames: the first (total % world) ranks get one more."""
    assert 0 <= rank < world and total_frames >= 0
    base, extra = divmod(total_frames, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def stream_owner(stream_id: int, world: int, streams_total: int):
    """Rank that owns a sensor stream under the strong (contiguous) split: consecutive frames of one stream must stay
    on one GPU because of the temporal state (adaptive thresholds / sensor height, reference patchworkpp.cpp:338-375)."""
    for r in range(world):
        if stream_id in strong_shard(streams_total, r, world):
            return r
    raise ValueError(stream_id)


class Dist:
    """Thin wrapper: works for world == 1 without initialising a process group."""

    def __init__(self, backend: str = None, device=None):
        self.rank, self.world, self.local = env_rank_world()
        self.device = device
        self.pg = False
        if self.world > 1:
            import torch.distributed as dist
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", self.local)
            if not dist.is_initialized():
                dist.init_process_group(backend, **kw)
            self.pg = True
            self.backend = backend

    def barrier(self):
        if self.pg:
            import torch.distributed as dist
            if self.backend == "nccl":
                dist.barrier(device_ids=[self.local])
            else:
                dist.barrier()

    def max_over_ranks(self, value: float) -> float:
        if not self.pg:
            return float(value)
        import torch.distributed as dist
        dev = torch.device("cuda", self.local) if self.backend == "nccl" else torch.device("cpu")
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_floats(self, value: float):
        """Every rank's value, in rank order."""
        if not self.pg:
            return [float(value)]
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, float(value))
        return out

    def gather_ints(self, values):
        """All ranks' integer lists concatenated in rank order (rank 0's view; used by tests for result checks)."""
        if not self.pg:
            return list(values)
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, list(values))
        return [v for part in out for v in part]

    def close(self):
        if self.pg:
            import torch.distributed as dist
            dist.destroy_process_group()
            self.pg = False
