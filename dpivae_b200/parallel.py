"""Batch-sharded data parallelism for the fused step (the reference has no
distributed code).  One process per GPU; rank r owns a shard of the rows of the global minibatch
-- a contiguous block, or (preferred with in-kernel noise) the CYCLIC shard r, r + G, r + 2G, ... --;
noise is indexed by GLOBAL row and the loss normalised by the GLOBAL batch inside the kernels, so
the only exchange is ONE allreduce(sum) of the flat [gradients | 8 loss scalars] buffer between
the backward and the (replicated, identical) fused Adam update."""
import torch
import torch.distributed as dist


def shard_bounds(n_rows, world_size, rank):
    """Contiguous row block of `rank` (first `n_rows % world_size` ranks get one extra row)."""
    base, rem = divmod(int(n_rows), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def torch_normal_grid_threads(numel, sm_count=148, max_threads_per_sm=2048):
    """Threads of the grid torch's CUDA `normal_` kernel launches for a tensor of `numel` elements
    (ATen/native/cuda/DistributionTemplates.h; dpivae_philox_plan in csrc/api.cu computes the same): generator thread
    idx draws the elements idx + grid_threads * (4 j + k), k = 0..3, of call j from ONE Philox evaluation."""
    grid = min((int(numel) + 255) // 256, int(sm_count) * (int(max_threads_per_sm) // 256))
    return 256 * max(grid, 1)


def cyclic_shards_own_whole_evaluations(grid_threads, nz, world_size, rows_global):
    """True when, with cyclic row shards (rank k owns global rows k, k + world, ...), the four elements of every Philox
    evaluation of the (n_mc, rows_global, nz) noise tensor belong to ONE rank -- the condition under which every rank
    can draw its noise with one evaluation per four elements (lat_noise_fill_cyclic_kernel; run_loss in csrc/api.cu
    applies the same test).  The elements are grid_threads apart in the flattened tensor = grid_threads / nz rows."""
    s, nz = int(world_size), int(nz)
    return s >= 1 and int(grid_threads) % (nz * s) == 0 and int(rows_global) % s == 0


def cyclic_local_slot(li, nz, world_size, rank, rows_global):
    """(owner rank, local flat index) of element `li` of the flattened (n_mc, rows_global, nz) noise tensor under
    cyclic row shards: the destination the cyclic noise pre-pass writes (local (m, row, i) order of the shard)."""
    nz, s, bg = int(nz), int(world_size), int(rows_global)
    il, trow = li % nz, li // nz
    m, grow = divmod(trow, bg)
    owner, lrow = grow % s, grow // s
    return owner, (m * (bg // s) + lrow) * nz + il


def allreduce_flat(buf, group=None):
    """Sum-allreduce of the flat gradient+scalar buffer (NCCL on GPUs, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


class DataParallelStep:
    """train step = local fused fwd+bwd on this rank's shard -> allreduce -> fused Adam."""

    def __init__(self, vae, groups, group=None):
        self.vae = vae
        self.eng = vae.engine()
        self.eng.set_groups(groups)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 1:
            # identical initial parameters everywhere
            dist.broadcast(self.eng.params, src=0, group=group)

    def step(self, x, c, y, n_mc, weights, B_global, row_offset, step, idx=None, max_grad_norm=0.0, row_stride=1):
        """`row_offset`, `row_stride`: global row of this rank's local row r is row_offset + r * row_stride.  Contiguous
        shards: (rank * rows, 1).  CYCLIC shards -- (rank, world) -- are the faster choice with in-kernel noise: each rank
        then owns whole Philox evaluations of torch's noise stream (include/dpivae_b200.h, dpivae_batch_t.row_stride)."""
        eng = self.eng
        if self.world == 1:
            eng.loss(x, c, y, n_mc, weights, True, idx=idx, B_global=B_global, row_offset=row_offset, adam_step=step,
                     max_grad_norm=max_grad_norm, row_stride=row_stride)
        else:
            eng.loss(x, c, y, n_mc, weights, True, idx=idx, B_global=B_global, row_offset=row_offset, row_stride=row_stride)
            allreduce_flat(eng.gradbuf, self.group)
            eng.adam_step(step, max_grad_norm)
        return eng.scalars
