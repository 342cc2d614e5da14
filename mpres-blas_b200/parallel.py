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


class LeanBroadcast:
    """Replication of B for the row-sharded mp_gemm that moves only what the chosen path reads.

    The exact-window fast path with the small-modulus base reads, per entry of B, its first n_in residues, its sign,
    its exponent and the UPPER interval bound: 4 n_in + 24 of the 4 N + 40 bytes (40 of 168 at the 424-bit set).  The
    receiving ranks therefore never hold a complete copy of B -- which is only sound while (a) the small base is the
    one chosen, (b) no entry falls back to the reference-order recomputation (that reads whole records) and (c) no
    rank's input conversion reads more than n_in residues.  `verify` checks exactly that on the device after the call
    (mpres_last_small_base, mpres_last_fallback_count); when it fails the caller repeats the step with the full
    broadcast (`broadcast_arrays`).  n_in is agreed once: the maximum over ranks of what a first, fully replicated
    call used.
    """

    def __init__(self, dist, N, src=0):
        self.dist, self.N, self.src, self.nin, self.buf = dist, N, src, 0, None

    def agree(self, nin_local, device):
        t = torch.tensor([int(nin_local)], dtype=torch.int32, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        self.nin = int(t.item())
        return self.nin

    def nbytes(self, count):
        return count * (4 * self.nin + 4 + 4 + 16)

    def broadcast(self, digits, sign, exp, ev):
        """digits [len * N] int32, sign / exp [len] int32, ev [2 * len * 2] int64 (lower bounds, then upper bounds)"""
        n = sign.numel()
        d2 = digits.view(n, self.N)
        if self.buf is None or self.buf.shape != (n, self.nin):
            self.buf = torch.empty((n, self.nin), dtype=digits.dtype, device=digits.device)
        rank = self.dist.get_rank()
        if rank == self.src:
            self.buf.copy_(d2[:, : self.nin])
        self.dist.broadcast(self.buf, src=self.src)
        if rank != self.src:
            d2[:, : self.nin].copy_(self.buf)
        self.dist.broadcast(sign, src=self.src)
        self.dist.broadcast(exp, src=self.src)
        self.dist.broadcast(ev[2 * n:], src=self.src)

    def verify(self, small_moduli, nin_used, fallback_count):
        return small_moduli > 0 and 0 < nin_used <= self.nin and fallback_count == 0


def soa_ranges(digits, sign, exp, ev, N, length, offset, count):
    """The five contiguous ranges of an mp_array_t's SoA arrays that hold elements offset .. offset + count: digits, sign, exp, and
    the lower / upper interval bounds (`ev`: int64 view of er_float_t[2 * length], lower bounds first -- src/types.cuh:85-92)."""
    return [digits[N * offset:N * (offset + count)], sign[offset:offset + count], exp[offset:offset + count],
            ev[2 * offset:2 * (offset + count)], ev[2 * (length + offset):2 * (length + offset + count)]]


def gather_column_shards(dist, full_ranges, my_ranges):
    """End-to-end recipe for the row-sharded GEMM: rank r has uploaded columns [n r / G, n (r + 1) / G) of B (elements
    k n r / G .. of the column-major array) from host memory; every rank ends up with all of B.  One all-gather per SoA range (the
    shards are equally long: n % G == 0), each rank contributing a copy of its own range."""
    for full, mine in zip(full_ranges, my_ranges):
        dist.all_gather_into_tensor(full, mine.clone())


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
