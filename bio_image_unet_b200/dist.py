"""One-process-per-GPU sharding helpers (torch.distributed over NCCL on the GPU box, gloo in CPU tests).

The tiled prediction path shards without any data-path collective: frames (2D), consecutive-frame pairs (Siam),
z-rows of the patch grid (3D) and volumes (multi-output 3D) are independent. Communication is only
(i) one all-reduce of the 65 536-bin intensity histogram ('first' / 'all' normalisation, the 3D global percentiles),
(ii) for overlapping 3D patches, point-to-point sends of the uint8 result patches of the z-rows that reach into a
neighbour's output slab, and (iii) the gather of the stitched output slabs on rank 0. (ii) and (iii) move DEVICE
tensors: NCCL send / recv straight into the destination buffer (ragged slabs need no padding), one D2H copy on rank
0 at the very end. With the gloo backend (CPU tests, or two ranks sharing one GPU) the same calls bounce through
host memory.
"""
import os
import warnings

import torch

from . import tiling


class DistContext:
    def __init__(self, enabled):
        ready = torch.distributed.is_available() and torch.distributed.is_initialized()
        if enabled and not ready:
            if int(os.environ.get('WORLD_SIZE', '1')) > 1:
                # under torchrun with init_process_group forgotten every rank would predict the whole stack and all
                # of them would write the same result file
                raise RuntimeError('distributed=True under a multi-process launch (WORLD_SIZE > 1), but '
                                   'torch.distributed is not initialised: call torch.distributed.init_process_group '
                                   'first')
            warnings.warn('distributed=True but torch.distributed is not initialised: running single-process',
                          RuntimeWarning, stacklevel=3)
        self.active = bool(enabled) and ready
        self.rank = torch.distributed.get_rank() if self.active else 0
        self.world = torch.distributed.get_world_size() if self.active else 1
        self.backend = torch.distributed.get_backend() if self.active else None

    @property
    def multi(self):
        return self.active and self.world > 1

    def device(self):
        n_dev = max(torch.cuda.device_count(), 1)      # ranks may share a device (gloo test runs on a 1-GPU box)
        return torch.device('cuda', int(os.environ.get('LOCAL_RANK', self.rank)) % n_dev)

    def shard(self, n_items):
        return tiling.shard_range(n_items, self.rank, self.world)

    def shards(self, n_items):
        return [tiling.shard_range(n_items, r, self.world) for r in range(self.world)]

    # ---- tensors as the backend wants them -------------------------------------------------------------------
    def _wire(self, t):
        """The tensor a collective can take: device tensors for NCCL, host tensors for gloo."""
        return t.cpu() if (self.backend == 'gloo' and t.is_cuda) else t

    def all_reduce_sum(self, t):
        if self.multi:
            w = self._wire(t)
            torch.distributed.all_reduce(w)
            if w is not t:
                t.copy_(w)
        return t

    def broadcast(self, t, src=0):
        if self.multi:
            w = self._wire(t)
            torch.distributed.broadcast(w, src)
            if w is not t:
                t.copy_(w)
        return t

    def barrier(self):
        if self.multi:
            torch.distributed.barrier()

    def exchange(self, sends, recvs):
        """Point-to-point exchange of contiguous tensors. sends: [(dst_rank, tensor)], recvs: [(src_rank, tensor)]
        (receive buffers are written in place). Pairs between two ranks match in list order."""
        if not self.multi or (not sends and not recvs):
            return
        ops, back = [], []
        for dst, t in sends:
            ops.append(torch.distributed.P2POp(torch.distributed.isend, self._wire(t.contiguous()), dst))
        for src, t in recvs:
            assert t.is_contiguous()
            w = self._wire(t)
            if w is not t:
                back.append((t, w))
            ops.append(torch.distributed.P2POp(torch.distributed.irecv, w, src))
        for req in torch.distributed.batch_isend_irecv(ops):
            req.wait()
        for t, w in back:
            t.copy_(w)

    def gather_slabs(self, local, bounds, dst=0, out=None, n_total=None):
        """Gather of ragged slabs along dim 0. `local`: this rank's (n_local, ...) tensor (device tensor under NCCL);
        bounds: [(lo, hi)] of every rank. Returns the full (n_total, ...) tensor on `dst` (None elsewhere); every
        slab is received directly into its place - no padding, no host bounce."""
        if not self.multi:
            return local
        n_total = max(b for _, b in bounds) if n_total is None else n_total
        full = None
        sends, recvs = [], []
        if self.rank == dst:
            full = out if out is not None else torch.empty((n_total, *local.shape[1:]), dtype=local.dtype,
                                                           device=local.device)
            lo, hi = bounds[dst]
            if hi > lo:
                full[lo:hi].copy_(local)
            recvs = [(r, full[a:b]) for r, (a, b) in enumerate(bounds) if r != dst and b > a]
        elif local.shape[0] > 0:
            sends = [(dst, local)]
        self.exchange(sends, recvs)
        return full

    def gather_frames(self, local, n_total, device=None):
        """numpy / tensor front end of gather_slabs for contiguous shard_range() slices; returns a numpy array on
        rank 0 (None elsewhere); identity when not distributed."""
        import numpy as np
        if not self.multi:
            return local
        t = torch.from_numpy(np.ascontiguousarray(local)) if isinstance(local, np.ndarray) else local
        if self.backend == 'nccl' and not t.is_cuda:
            t = t.to(device if device is not None else self.device())
        full = self.gather_slabs(t, self.shards(n_total))
        return None if full is None else full.cpu().numpy()
