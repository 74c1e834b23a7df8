"""Samples-per-pixel split across the GPUs of one node.

Iterations are independent: the RNG depends only on (iter, pixelIndex, depth) (src/pathtrace.cu:41-45)
and the only cross-iteration state is the additive accumulation buffer (src/pathtrace.cu:276).  Rank r
of G renders iterations r+1, r+1+G, ... into a private W*H*3 float buffer; ONE sum-reduce of that
buffer to rank 0 per output frame is the only collective (NCCL over NVLink/NVSwitch on GPUs, gloo in
the CPU tests).  One process per GPU; torch.distributed is plumbing only.
"""


def rank_iterations(first_iter, total_iters, rank, world):
    """(iterBegin, count, iterStride) for ptb_render on `rank`: the interleaved split keeps every
    prefix of the sample sequence balanced.  The union over ranks is exactly
    first_iter .. first_iter + total_iters - 1, each iteration once."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    count = (total_iters - rank + world - 1) // world if total_iters > rank else 0
    return first_iter + rank, count, world


def reduce_frame(accum, frame=None, dst=0, group=None):
    """Sum the per-rank accumulation buffers into `frame` on rank `dst` (a copy first: the private
    partial sums keep growing across frames, so they must not be overwritten by the total)."""
    import torch.distributed as dist
    if frame is None:
        frame = accum.clone()
    else:
        frame.copy_(accum)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(frame, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return frame


def init_library_comm(renderer, rank, world, pkg):
    """One process per GPU (torchrun): give `renderer`'s context an NCCL communicator INSIDE libptb200.so (ptb_comm_init_rank).
    Rank 0 asks the library for the ncclUniqueId; torch.distributed only carries those 128 bytes to the other ranks."""
    import torch.distributed as dist
    box = [pkg.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    renderer.comm_init_rank(box[0], rank, world)


def reduce_frame_library(renderer, root=0):
    """The frame collective through the library: ONE ncclReduce(sum, float32, 3*W*H) of every rank's private accumulation
    buffer into the root's frame buffer, on the renderer's own stream (include/ptb200.h ptb_comm_reduce)."""
    renderer.comm_reduce(root)
