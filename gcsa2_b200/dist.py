"""Multi-GPU plumbing: queries shard, the index is replicated, counters are gathered.

The reference has no distributed mode; its callers parallelise over independent queries with
OpenMP (src/algorithms.cpp:113, 409).  The same independence carries over: rank r of W searches
the contiguous block [r * n / W, (r + 1) * n / W) of the batch against its own replica of the
index, there is no exchange during the search, and one all-reduce sums the result counters
(torch.distributed: NCCL over NVLink on the GPUs, gloo in the CPU tests).
"""
import numpy as np

COUNTER_NAMES = ("queries", "found", "total_length", "occurrences")


def shard_bounds(n, rank, world):
    """Contiguous block of a batch of n queries owned by `rank` (balanced to within one query)."""
    base, extra = divmod(int(n), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_patterns(chars, offsets, rank, world):
    """Slice (chars, offsets) to the rank's block; offsets are rebased to start at 0."""
    n = len(offsets) - 1
    q0, q1 = shard_bounds(n, rank, world)
    c0, c1 = int(offsets[q0]), int(offsets[q1])
    return chars[c0:c1], (offsets[q0:q1 + 1] - offsets[q0]).astype(np.uint64), (q0, q1)


def shard_patterns_by_length(chars, offsets, rank, world):
    """Shard of a batch of patterns of very different lengths (BASELINE.json configs[4]: 16-256 bp): the patterns
    are ordered by length and dealt round-robin, so every rank gets the same mix of lengths and the same number of
    characters to within one pattern per length class.  Returns (chars, offsets, ids): the rank's patterns packed
    contiguously in increasing id order, and their indexes in the original batch (for scattering the results back)."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    lengths = np.diff(offsets.astype(np.int64))
    order = np.argsort(lengths, kind="stable")
    ids = np.sort(order[int(rank)::int(world)])
    mine = lengths[ids]
    out_offsets = np.zeros(ids.size + 1, dtype=np.uint64)
    out_offsets[1:] = np.cumsum(mine)
    # gather the characters of the selected patterns: positions offsets[id] + (0 .. length - 1)
    starts = np.repeat(offsets[ids].astype(np.int64) - out_offsets[:-1].astype(np.int64), mine)
    index = starts + np.arange(int(out_offsets[-1]), dtype=np.int64)
    out_chars = np.asarray(chars)[index] if index.size else np.zeros(1, dtype=np.uint8)
    return out_chars, out_offsets, ids


def find_counters(sp, ep, occurrences=0):
    """Per-shard result counters of a find() batch (what a caller aggregates: query_gcsa.cpp:98-103)."""
    sp = np.asarray(sp, dtype=np.uint64); ep = np.asarray(ep, dtype=np.uint64)
    nonempty = (sp + np.uint64(1)) <= (ep + np.uint64(1))
    length = int((ep[nonempty] + np.uint64(1) - sp[nonempty]).sum()) if nonempty.any() else 0
    return np.array([sp.size, int(nonempty.sum()), length, int(occurrences)], dtype=np.int64)


def all_reduce_counters(counters, device=None):
    """Sum of the counter vectors over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(counters, dtype=np.int64)
    t = torch.as_tensor(np.asarray(counters, dtype=np.int64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def gather_ranges(sp, ep, device=None):
    """All-gather of equal-sized result blocks (callers that want every range on every rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(sp), np.asarray(ep)
    world = dist.get_world_size()
    local = torch.as_tensor(np.stack([np.asarray(sp, dtype=np.uint64).view(np.int64), np.asarray(ep, dtype=np.uint64).view(np.int64)]))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    mine = torch.tensor([local.shape[1]], dtype=torch.int64)
    if device is not None:
        local, mine, sizes = local.to(device), mine.to(device), [s.to(device) for s in sizes]
    dist.all_gather(sizes, mine)
    m = int(max(int(s) for s in sizes))
    padded = torch.zeros((2, m), dtype=torch.int64, device=local.device); padded[:, :local.shape[1]] = local
    blocks = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(blocks, padded)
    out = torch.cat([b[:, :int(s)] for b, s in zip(blocks, sizes)], dim=1).cpu().numpy()
    return out[0].view(np.uint64), out[1].view(np.uint64)
