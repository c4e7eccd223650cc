"""Data-parallel contract on CPU (gloo, world_size 2): shards of one global minibatch, noise indexed by
global row, loss normalised by the global batch, ONE allreduce(sum) of the flat [grads | scalars] buffer
== the single-process result.  The per-shard arithmetic here is the oracle (checker) -- the CUDA side of
the same contract is tests/test_gpu_parity.py::test_shard_invariance_and_determinism."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path, cyclic=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as gu
    from dpivae_b200.parallel import allreduce_flat, shard_bounds
    from oracle import dpivae_oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    g, spec, sd = gu.load("bridge", "P")
    spec = orc.cast_spec(spec, torch.float64)
    sd = {k: v.double() for k, v in sd.items()}
    x, c, y = (torch.from_numpy(g[k]).double() for k in "xcy")
    B = x.shape[0]
    eps = tuple(e.double() for e in gu.eps_of(g, spec))
    if cyclic:   # rank k owns the global rows k, k + world, ... (dpivae_batch_t.row_stride; DataParallelStep's default)
        rows = slice(rank, None, world)
    else:
        lo, hi = shard_bounds(B, world, rank)
        rows = slice(lo, hi)
    eps_s = tuple(e[:, rows] for e in eps)
    scal, _, _, grads = orc.loss_and_grads(sd, spec, x[rows], c[rows], y[rows], eps_s, n_batch=B)
    flat = torch.cat([grads[k].reshape(-1) for k in spec["trainable"]] + [torch.stack(scal)])
    allreduce_flat(flat)
    if rank == 0:
        torch.save(flat, out_path)
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("cyclic", [False, True])
def test_two_shards_allreduce_equals_full_batch(tmp_path, cyclic):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as gu
    from oracle import dpivae_oracle as orc

    out = str(tmp_path / "flat.pt")
    port = 29500 + (os.getpid() % 2000) + (7 if cyclic else 0)
    mp.spawn(_worker, args=(2, port, out, cyclic), nprocs=2, join=True)
    flat = torch.load(out)
    g, spec, sd = gu.load("bridge", "P")
    spec = orc.cast_spec(spec, torch.float64)
    sd = {k: v.double() for k, v in sd.items()}
    x, c, y = (torch.from_numpy(g[k]).double() for k in "xcy")
    eps = tuple(e.double() for e in gu.eps_of(g, spec))
    scal, _, _, grads = orc.loss_and_grads(sd, spec, x, c, y, eps)
    ref = torch.cat([grads[k].reshape(-1) for k in spec["trainable"]] + [torch.stack(scal)])
    assert gu.rel_l2(flat, ref) < 1e-12


def test_shard_bounds_cover_rows():
    from dpivae_b200.parallel import shard_bounds

    for n in (1, 7, 64, 1000, 1048576):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
