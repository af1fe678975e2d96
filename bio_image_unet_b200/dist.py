"""One-process-per-GPU sharding helpers (torch.distributed over NCCL on the GPU box, gloo in CPU tests).

The tiled prediction path shards without any data-path collective: frames (2D), consecutive-frame pairs (Siam),
z-rows of the patch grid (3D) and volumes (multi-output 3D) are independent. Collectives are
only used for (i) the stack-wide intensity histogram of the 'first' / 'all' normalisation modes and the 3D global
percentiles (one 256 KiB all-reduce) and (ii) gathering the stitched uint8 output on rank 0.
"""
import numpy as np
import torch

from . import tiling


class DistContext:
    def __init__(self, enabled):
        self.active = bool(enabled) and torch.distributed.is_available() and torch.distributed.is_initialized()
        self.rank = torch.distributed.get_rank() if self.active else 0
        self.world = torch.distributed.get_world_size() if self.active else 1

    def device(self):
        import os
        return torch.device('cuda', int(os.environ.get('LOCAL_RANK', self.rank % max(torch.cuda.device_count(), 1))))

    def shard(self, n_items):
        return tiling.shard_range(n_items, self.rank, self.world)

    def all_reduce_sum(self, t):
        if self.active and self.world > 1:
            if torch.distributed.get_backend() == 'gloo' and t.is_cuda:
                c = t.cpu()
                torch.distributed.all_reduce(c)
                return c.to(t.device)
            torch.distributed.all_reduce(t)
        return t

    def broadcast(self, t, src=0):
        if self.active and self.world > 1:
            torch.distributed.broadcast(t, src)
        return t

    def gather_frames(self, local, n_total, device):
        """local: (n_local, ...) uint8/float32 ndarray of this rank's contiguous slice. Returns the full array on
        rank 0 (None elsewhere); identity when not distributed."""
        if not (self.active and self.world > 1):
            return local
        backend = torch.distributed.get_backend()
        counts = [tiling.shard_range(n_total, r, self.world) for r in range(self.world)]
        max_n = max(b - a for a, b in counts)
        pad = np.zeros((max_n, *local.shape[1:]), dtype=local.dtype)
        pad[:local.shape[0]] = local
        t = torch.from_numpy(pad)
        if backend == 'nccl':
            t = t.to(device)
        bufs = [torch.empty_like(t) for _ in range(self.world)] if self.rank == 0 else None
        torch.distributed.gather(t, bufs, dst=0)
        if self.rank != 0:
            return None
        parts = [bufs[r][:b - a].cpu().numpy() for r, (a, b) in enumerate(counts)]
        return np.concatenate(parts, axis=0)
