"""Batch-sharded data parallelism for the fused step (the reference has no
distributed code).  One process per GPU; rank r owns rows [r*B/G, (r+1)*B/G) of the global
minibatch; noise is indexed by GLOBAL row and the loss normalised by the GLOBAL batch inside the
kernels, so the only exchange is ONE allreduce(sum) of the flat [gradients | 8 loss scalars]
buffer between the backward and the (replicated, identical) fused Adam update."""
import torch
import torch.distributed as dist


def shard_bounds(n_rows, world_size, rank):
    """Contiguous row block of `rank` (first `n_rows % world_size` ranks get one extra row)."""
    base, rem = divmod(int(n_rows), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


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
