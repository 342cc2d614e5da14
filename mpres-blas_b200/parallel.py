"""Multi-GPU host logic for the path (one process per GPU, torch.distributed for the plumbing).

The reference has no multi-GPU code; the partitioning follows SURVEY 8(e):
  * mp_gemm  -- A and C split into row blocks (each rank holds compact shards, lda = ldc = rows),
                B replicated by a broadcast of its four SoA arrays; no reduction.
  * mp_gemv  -- trans = N: row blocks of A and y, x replicated; trans = T: partial y per rank, gathered
                and summed in RNS.
  * mp_dot   -- contiguous vector segments; every rank reduces its segment to ONE packed mp_float_t
                (4N + 40 bytes), the partials are all-gathered as bytes and summed in rank order with
                mp_add on every rank (NCCL cannot sum this type), so all ranks hold the same bits.
The numeric work is injected as callables so that the same logic runs over the CUDA library (NCCL) and,
in the CPU tests, over stand-ins (gloo).
"""
import torch


def row_block(m, world, rank):
    """rows [lo, hi) of an m-row matrix owned by `rank`"""
    return (m * rank) // world, (m * (rank + 1)) // world


def segment(n, world, rank):
    return (n * rank) // world, (n * (rank + 1)) // world


def broadcast_arrays(dist, tensors, src=0):
    """replicate the SoA arrays of one mp_array_t (digits, sign, exp, eval)"""
    for t in tensors:
        dist.broadcast(t, src=src)


def gemm_row_sharded(dist, b_tensors, local_gemm, src=0):
    """C_r = alpha * A_r * B + beta * C_r on every rank r after replicating B"""
    if dist is not None and dist.get_world_size() > 1:
        broadcast_arrays(dist, b_tensors, src)
    return local_gemm()


def dot_segment_sharded(dist, local_partial, reduce_partials):
    """local_partial() -> uint8 tensor holding one packed mp_float_t; reduce_partials(bytes, count) sums
    `count` packed records in index order and returns whatever the caller's result type is"""
    part = local_partial()
    if dist is None or dist.get_world_size() == 1:
        return reduce_partials(part, 1)
    world = dist.get_world_size()
    gathered = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(gathered, part)
    return reduce_partials(torch.cat(gathered), world)


def gemv_t_sharded(dist, local_partial_y, reduce_columns):
    """trans = T with row-sharded A and x: every rank produces a full-length partial y (packed records);
    all-gather, then element j of y = sum over ranks in rank order"""
    part = local_partial_y()
    if dist is None or dist.get_world_size() == 1:
        return reduce_columns([part])
    world = dist.get_world_size()
    gathered = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(gathered, part)
    return reduce_columns(gathered)
