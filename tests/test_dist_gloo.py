"""world_size-2 gloo tests (CPU) of the sharding logic used on the GPU box with NCCL: contiguous shard ranges,
histogram all-reduce, gather of the per-rank result slices, Siam pair bookkeeping across a shard boundary."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from bio_image_unet_b200.dist import DistContext
        ctx = DistContext(True)
        assert ctx.active and ctx.rank == rank and ctx.world == world
        # frames of a 7-frame movie: rank 0 -> [0, 4), rank 1 -> [4, 7)
        lo, hi = ctx.shard(7)
        assert (lo, hi) == ((0, 4) if rank == 0 else (4, 7))
        # stack-wide histogram = sum of the per-rank partial histograms
        rng = np.random.default_rng(0)
        movie = rng.integers(0, 500, (7, 16, 16)).astype('uint16')
        part = torch.from_numpy(np.bincount(movie[lo:hi].ravel(), minlength=65536).astype('int32'))[None]
        total = ctx.all_reduce_sum(part.clone())
        assert np.array_equal(total[0].numpy(), np.bincount(movie.ravel(), minlength=65536))
        # gather of uneven result slices on rank 0
        local = (movie[lo:hi] % 251).astype('uint8')[:, None]
        full = ctx.gather_frames(local, 7, torch.device('cpu'))
        if rank == 0:
            assert np.array_equal(full, (movie % 251).astype('uint8')[:, None])
        else:
            assert full is None
        # float32 volumes (multi-output 3D): 3 volumes over 2 ranks
        vlo, vhi = ctx.shard(3)
        vols = np.arange(3 * 2 * 4, dtype='float32').reshape(3, 2, 4)
        fullv = ctx.gather_frames(vols[vlo:vhi], 3, torch.device('cpu'))
        if rank == 0:
            assert np.array_equal(fullv, vols)
        # ragged slabs straight into place (3D: output planes owned per rank; trailing ranks may own nothing)
        slabs = [(0, 5), (5, 7)]
        vol = torch.arange(7 * 3 * 2, dtype=torch.uint8).reshape(7, 3, 2)
        got = ctx.gather_slabs(vol[slabs[rank][0]:slabs[rank][1]].clone(), slabs, n_total=7)
        assert (got is None) if rank else torch.equal(got, vol)
        got = ctx.gather_slabs(vol[:7 if rank == 0 else 0].clone(), [(0, 7), (0, 0)], n_total=7)
        assert (got is None) if rank else torch.equal(got, vol)
        # point-to-point exchange (3D boundary patches travel from the higher to the lower rank)
        if rank == 1:
            ctx.exchange([(0, torch.full((4, 2), 9, dtype=torch.uint8))], [])
        else:
            buf = torch.zeros((4, 2), dtype=torch.uint8)
            ctx.exchange([], [(1, buf)])
            assert int(buf.sum()) == 72
        # Siam: a rank's first pair needs the last frame of the previous rank (read from the input, no exchange)
        prev = [(1 if i == 0 else i - 1) for i in range(lo, hi)]
        assert prev == ([1, 0, 1, 2] if rank == 0 else [3, 4, 5])
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_distributed_without_process_group_is_loud(monkeypatch):
    """distributed=True with no process group: a warning single-process, an error under a multi-process launch (every
    rank would otherwise predict the whole stack and write the same file)."""
    import pytest
    from bio_image_unet_b200.dist import DistContext
    monkeypatch.delenv('WORLD_SIZE', raising=False)
    with pytest.warns(RuntimeWarning):
        ctx = DistContext(True)
    assert not ctx.active and ctx.world == 1
    monkeypatch.setenv('WORLD_SIZE', '2')
    with pytest.raises(RuntimeError):
        DistContext(True)


def test_two_rank_sharding_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / 'ok0') and os.path.exists(tmp_path / 'ok1')
