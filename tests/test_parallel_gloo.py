"""CPU, world_size 2 over gloo: the data-parallel exchange step.  Rank-local gradient of the
local batch mean, summed by all-reduce and scaled by 1/world, must equal the single-process
gradient of the concatenated batch, and the shard arithmetic must tile
the batch exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parallel


def test_shard_range_tiles_the_batch():
    for total in (1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmp):
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (os.path.join(root, 'ssd-tensorflow_b200'), os.path.join(root, 'oracle')):
        sys.path.insert(0, p)
    import net_oracle as no
    import parallel as par
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    B, A = 4, 60
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, A, 16, generator=g, dtype=torch.float64)
    labels = torch.zeros(B, A, 25, dtype=torch.float64); labels[..., 20] = 1
    for b in range(B):
        for a in range(3 + b):
            labels[b, a * 7, 20] = 0; labels[b, a * 7, (a + b) % 20] = 1
            labels[b, a * 7, 21:] = torch.randn(4, generator=g, dtype=torch.float64)
    w = torch.randn(16, 25, generator=g, dtype=torch.float64, requires_grad=True)

    def loss_of(lo, hi):
        out = feats[lo:hi] @ w
        c, l = no.multibox_loss(out, labels[lo:hi])
        return c + l
    lo, hi = par.shard_range(B, rank, world)
    (gl,) = torch.autograd.grad(loss_of(lo, hi), w)
    gl = gl.clone()
    # once in one piece, once bucket by bucket in the order the engine's backward completes them (tail of the flat buffer
    # first): both must give the same averaged gradient
    flat = gl.reshape(-1).clone()
    n = flat.numel()
    buckets = [(2 * n // 3, n), (n // 3, 2 * n // 3), (0, n // 3)]
    seen = []
    scale_b = par.allreduce_buckets(flat, buckets, world, before=seen.append)
    assert seen == [0, 1, 2] and scale_b == 1.0 / world
    scale = par.average_gradients(gl, world)
    assert torch.equal(flat.reshape(gl.shape), gl)
    gl *= scale
    (gfull,) = torch.autograd.grad(loss_of(0, B), w)
    err = float((gl - gfull).abs().max() / gfull.abs().max())
    np.save(os.path.join(tmp, 'err%d.npy' % rank), np.array([err]))
    dist.destroy_process_group()


def test_allreduce_average_equals_full_batch_gradient(tmp_path):
    world = 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / ('err%d.npy' % r))[0] < 1e-12
