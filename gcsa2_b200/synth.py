"""Synthetic inputs of the benchmark configurations: linear references,
SNP-bubble variation graphs, and pattern batches sampled from them.  All generators are seeded
numpy Generators (PCG64), so host and test code regenerate identical data."""
import numpy as np

from .builder import CharGraph
from .flat import NODE_ID_OFFSET

SOURCE_COMP, SINK_COMP = 6, 0
COMP2CHAR = np.frombuffer(b"$ACGTN#", dtype=np.uint8)


def random_sequence(length, seed):
    """Uniform i.i.d. ACGT as comp values 1..4."""
    return np.random.default_rng(seed).integers(1, 5, size=int(length), dtype=np.uint8)


def _backbone_values(length, breaks, node_len, first_id):
    """vg-style ids for backbone positions: a new node starts at every break and every node_len
    characters; value = id << 11 | offset (include/gcsa/support.h:443-471)."""
    pos = np.arange(length, dtype=np.int64)
    seg_start = np.where(breaks, pos, 0)
    seg_start = np.maximum.accumulate(seg_start)
    rel = pos - seg_start
    new_node = (rel % node_len) == 0
    ids = np.cumsum(new_node) - 1 + first_id
    return (ids.astype(np.uint64) << np.uint64(NODE_ID_OFFSET)) | (rel % node_len).astype(np.uint64), int(ids[-1]) + 1


def linear_graph(seq, node_len=32):
    """# -> seq[0] -> ... -> seq[L-1] -> $ ; node 0 is the source, node L + 1 the sink."""
    L = int(seq.size)
    comp = np.empty(L + 2, dtype=np.uint8)
    comp[0], comp[1:L + 1], comp[L + 1] = SOURCE_COMP, seq, SINK_COMP
    breaks = np.zeros(L, dtype=bool); breaks[0] = True
    bvals, next_id = _backbone_values(L, breaks, node_len, first_id=2)
    value = np.empty(L + 2, dtype=np.uint64)
    value[0] = np.uint64(1 << NODE_ID_OFFSET); value[1:L + 1] = bvals
    value[L + 1] = np.uint64(next_id << NODE_ID_OFFSET)
    succ_offsets = np.minimum(np.arange(L + 3, dtype=np.uint64), np.uint64(L + 1))
    succ = np.arange(1, L + 2, dtype=np.uint64)
    return CharGraph(comp=comp, value=value, succ_offsets=succ_offsets, succ=succ, sink=L + 1,
                     sources=np.array([0], dtype=np.uint64))


def snp_graph(seq, seed, snp_rate=0.01, node_len=32):
    """Backbone seq with biallelic SNP bubbles at ~snp_rate of the positions (never adjacent,
    never at the first or last base).  Node numbering: 0 = source, 1..L = backbone (reference
    alleles), L+1..L+S = alternative alleles in site order, L+S+1 = sink.
    Returns (CharGraph, sites, alt_comps)."""
    rng = np.random.default_rng(seed)
    L = int(seq.size)
    cand = np.flatnonzero(rng.random(L) < snp_rate)
    cand = cand[(cand > 0) & (cand < L - 1)]
    if cand.size:
        keep = np.ones(cand.size, dtype=bool)
        keep[1:] = np.diff(cand) >= 2
        # drop the second of every adjacent pair (after the first pass no two kept are adjacent
        # unless three in a row: iterate once more)
        cand = cand[keep]
        keep = np.ones(cand.size, dtype=bool); keep[1:] = np.diff(cand) >= 2
        cand = cand[keep]
    sites = cand.astype(np.int64)
    S = int(sites.size)
    alt = ((seq[sites].astype(np.int64) - 1 + rng.integers(1, 4, size=S)) % 4 + 1).astype(np.uint8)

    n = L + S + 2
    sink = L + S + 1
    comp = np.empty(n, dtype=np.uint8)
    comp[0], comp[1:L + 1], comp[L + 1:L + 1 + S], comp[sink] = SOURCE_COMP, seq, alt, SINK_COMP

    is_site = np.zeros(L + 1, dtype=bool); is_site[sites] = True
    breaks = np.zeros(L, dtype=bool); breaks[0] = True
    breaks[sites] = True; breaks[sites + 1] = True
    bvals, next_id = _backbone_values(L, breaks, node_len, first_id=2)
    value = np.empty(n, dtype=np.uint64)
    value[0] = np.uint64(1 << NODE_ID_OFFSET); value[1:L + 1] = bvals
    value[L + 1:L + 1 + S] = (np.arange(next_id, next_id + S, dtype=np.uint64) << np.uint64(NODE_ID_OFFSET))
    value[sink] = np.uint64((next_id + S) << NODE_ID_OFFSET)

    # out-degree: node for backbone position p (index p + 1) and the source precede position p + 1
    # (or the sink); one extra successor if position p + 1 is a site.  Alt alleles have one successor.
    deg = np.ones(n, dtype=np.uint64); deg[sink] = 0
    site_index = np.full(L + 1, -1, dtype=np.int64); site_index[sites] = np.arange(S)
    deg[0:L] += is_site[0:L].astype(np.uint64)     # node index i (0..L-1) precedes position i
    succ_offsets = np.zeros(n + 1, dtype=np.uint64); succ_offsets[1:] = np.cumsum(deg)
    succ = np.empty(int(succ_offsets[-1]), dtype=np.uint64)
    # node i in 0..L-1 precedes position i (source precedes position 0, backbone p-1 precedes p)
    base = succ_offsets[0:L].astype(np.int64)
    succ[base] = np.arange(1, L + 1, dtype=np.uint64)                      # reference allele of position i
    with_alt = np.flatnonzero(is_site[0:L])
    succ[base[with_alt] + 1] = (L + 1 + site_index[with_alt]).astype(np.uint64)
    succ[int(succ_offsets[L])] = sink                                         # last base -> sink
    alt_nodes = np.arange(L + 1, L + 1 + S)
    succ[succ_offsets[alt_nodes].astype(np.int64)] = (sites + 2).astype(np.uint64)   # alt at p -> position p + 1
    graph = CharGraph(comp=comp, value=value, succ_offsets=succ_offsets, succ=succ, sink=sink,
                      sources=np.array([0], dtype=np.uint64))
    return graph, sites, alt


def patterns_from_sequence(seq, n, length, seed):
    """n substrings of the linear reference as ASCII bytes (all occur, `length` backward steps)."""
    rng = np.random.default_rng(seed)
    starts = rng.integers(0, seq.size - length + 1, size=int(n), dtype=np.int64)
    idx = starts[:, None] + np.arange(length, dtype=np.int64)[None, :]
    chars = COMP2CHAR[seq[idx]].reshape(-1)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
    return chars, offsets


def random_patterns(n, length, seed):
    """Uniform random ACGT strings (miss after ~log4(N) steps in a random reference)."""
    rng = np.random.default_rng(seed)
    chars = COMP2CHAR[rng.integers(1, 5, size=int(n) * int(length), dtype=np.uint8)]
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
    return chars, offsets


def patterns_from_snp_graph(seq, sites, alt, n, length, seed, error_rate=0.0):
    """Random walks through the SNP graph: a backbone window where each site takes the
    alternative allele with probability 1/2; optional substitutions."""
    rng = np.random.default_rng(seed)
    starts = rng.integers(0, seq.size - length + 1, size=int(n), dtype=np.int64)
    idx = starts[:, None] + np.arange(length, dtype=np.int64)[None, :]
    alt_full = np.zeros(seq.size, dtype=np.uint8); alt_full[sites] = alt
    comps = seq[idx]
    use_alt = (alt_full[idx] > 0) & (rng.random(idx.shape) < 0.5)
    comps = np.where(use_alt, alt_full[idx], comps)
    if error_rate > 0:
        err = rng.random(idx.shape) < error_rate
        comps = np.where(err, (comps - 1 + rng.integers(1, 4, size=idx.shape)) % 4 + 1, comps).astype(np.uint8)
    chars = COMP2CHAR[comps].reshape(-1)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
    return chars, offsets


def mixed_length_patterns(seq, sites, alt, n, min_len, max_len, seed, error_rate=0.01):
    """Config 5: lengths uniform in [min_len, max_len], sampled from the graph with substitutions."""
    rng = np.random.default_rng(seed)
    lengths = rng.integers(min_len, max_len + 1, size=int(n), dtype=np.int64)
    offsets = np.zeros(n + 1, dtype=np.uint64); offsets[1:] = np.cumsum(lengths)
    starts = rng.integers(0, seq.size - max_len, size=int(n), dtype=np.int64)
    total = int(offsets[-1])
    owner = np.repeat(np.arange(n), lengths)
    within = np.arange(total) - np.repeat(offsets[:-1].astype(np.int64), lengths)
    idx = starts[owner] + within
    alt_full = np.zeros(seq.size, dtype=np.uint8); alt_full[sites] = alt
    comps = seq[idx]
    use_alt = (alt_full[idx] > 0) & (rng.random(total) < 0.5)
    comps = np.where(use_alt, alt_full[idx], comps)
    err = rng.random(total) < error_rate
    comps = np.where(err, (comps - 1 + rng.integers(1, 4, size=total)) % 4 + 1, comps).astype(np.uint8)
    return COMP2CHAR[comps], offsets


# ---- device-side generators (configs[3]: a 3 Gbp reference and 125 M patterns per GPU never touch the host) -------------
# A counter-based generator (splitmix64 of the element index) so that every rank, and the host, regenerate the same data.

_SM_GOLDEN, _SM_C1, _SM_C2 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB


def _as_i64(x):
    return x - (1 << 64) if x >= (1 << 63) else x


def splitmix64_numpy(index, seed):
    """splitmix64 of (index + (seed + 1) * golden), as uint64."""
    with np.errstate(over="ignore"):
        x = index.astype(np.uint64) + np.uint64((seed + 1) * _SM_GOLDEN & ((1 << 64) - 1))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(_SM_C1)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(_SM_C2)
        return x ^ (x >> np.uint64(31))


def _splitmix64_torch(index, seed):
    """The same function on an int64 CUDA tensor (two's complement arithmetic; logical shifts by masking)."""
    x = index + _as_i64((seed + 1) * _SM_GOLDEN & ((1 << 64) - 1))
    x = (x ^ ((x >> 30) & ((1 << 34) - 1))) * _as_i64(_SM_C1)
    x = (x ^ ((x >> 27) & ((1 << 37) - 1))) * _as_i64(_SM_C2)
    return x ^ ((x >> 31) & ((1 << 33) - 1))


def counter_sequence(length, seed):
    """Host version of device_sequence (tests)."""
    return (1 + (splitmix64_numpy(np.arange(int(length), dtype=np.uint64), seed) >> np.uint64(62))).astype(np.uint8)


def device_sequence(length, seed, device="cuda", chunk=1 << 27):
    """Uniform i.i.d. ACGT as comp values 1..4 in a CUDA uint8 tensor: 1 + the top two bits of splitmix64(i, seed)."""
    import torch
    out = torch.empty(int(length), dtype=torch.uint8, device=device)
    for a in range(0, int(length), chunk):
        b = min(int(length), a + chunk)
        h = _splitmix64_torch(torch.arange(a, b, dtype=torch.int64, device=device), seed)
        out[a:b] = (1 + ((h >> 62) & 3)).to(torch.uint8)
    return out


def device_pattern_starts(seq_length, n, length, seed, device="cuda"):
    import torch
    h = _splitmix64_torch(torch.arange(int(n), dtype=torch.int64, device=device), seed)
    return ((h >> 1) & ((1 << 62) - 1)) % (int(seq_length) - int(length) + 1)


def device_patterns(seq, n, length, seed, chunk=1 << 23):
    """n substrings of the device-resident reference as ASCII bytes (CUDA uint8 tensor of n * length bytes)."""
    import torch
    starts = device_pattern_starts(seq.numel(), n, length, seed, device=seq.device)
    windows = seq.unfold(0, int(length), 1)                   # (L - length + 1, length) view, no copy
    lut = torch.tensor(list(COMP2CHAR), dtype=torch.uint8, device=seq.device)
    out = torch.empty(int(n) * int(length), dtype=torch.uint8, device=seq.device)
    for a in range(0, int(n), chunk):
        b = min(int(n), a + chunk)
        comps = windows[starts[a:b]]
        out[a * length:b * length] = lut[comps.reshape(-1).long()]
    return out


def device_mixed_length_patterns(seq, sites, alt, n, min_len, max_len, seed, error_rate=0.01):
    """Config 5 on the device: n patterns, lengths uniform in [min_len, max_len], each a walk through the SNP graph
    (every site takes the alternative allele with probability 1/2) with substitutions at `error_rate`.  seq: CUDA uint8
    comps; sites / alt: host arrays of snp_graph().  Counter-based randomness (splitmix64 of the pattern / character
    index), so every rank regenerates the same batch.  -> (chars uint8 CUDA tensor of ASCII bytes, offsets int64 CUDA
    tensor of n + 1)."""
    import torch
    dev = seq.device
    L = int(seq.numel())
    ids = torch.arange(int(n), dtype=torch.int64, device=dev)
    h = _splitmix64_torch(ids, seed)
    lengths = int(min_len) + (((h >> 1) & ((1 << 62) - 1)) % (int(max_len) - int(min_len) + 1))
    starts = ((_splitmix64_torch(ids, seed + 1) >> 1) & ((1 << 62) - 1)) % (L - int(max_len))
    offsets = torch.zeros(int(n) + 1, dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(lengths, 0)
    total = int(offsets[-1].item())
    alt_full = torch.zeros(L, dtype=torch.uint8, device=dev)
    alt_full[torch.as_tensor(np.asarray(sites, dtype=np.int64), device=dev)] = torch.as_tensor(np.asarray(alt, dtype=np.uint8), device=dev)
    lut = torch.tensor(list(COMP2CHAR), dtype=torch.uint8, device=dev)
    chars = torch.empty(total, dtype=torch.uint8, device=dev)
    block = 1 << 18                                                   # patterns per block (bounds the int64 temporaries)
    for a in range(0, int(n), block):
        b = min(int(n), a + block)
        c0, c1 = int(offsets[a].item()), int(offsets[b].item())
        owner = torch.repeat_interleave(torch.arange(a, b, dtype=torch.int64, device=dev), lengths[a:b])
        k = torch.arange(c0, c1, dtype=torch.int64, device=dev)
        pos = starts[owner] + (k - offsets[owner])
        comps = seq[pos]
        r = _splitmix64_torch(k, seed + 2)
        alts = alt_full[pos]
        use_alt = (alts > 0) & (((r >> 40) & 1) == 1)
        comps = torch.where(use_alt, alts, comps)
        err = ((r >> 8) & 0xFFFFFF) < int(error_rate * (1 << 24))
        shifted = ((comps.to(torch.int64) - 1 + 1 + ((r >> 34) & 3) % 3) % 4 + 1).to(torch.uint8)
        comps = torch.where(err, shifted, comps)
        chars[c0:c1] = lut[comps.long()]
    return chars, offsets


def device_shard_by_length(chars, offsets, rank, world):
    """dist.shard_patterns_by_length on the device: the patterns ordered by length are dealt round-robin, the rank's
    share is packed contiguously in increasing id order.  -> (chars, offsets, ids) CUDA tensors."""
    import torch
    lengths = offsets[1:] - offsets[:-1]
    order = torch.sort(lengths, stable=True).indices
    ids = torch.sort(order[int(rank)::int(world)]).values
    mine = lengths[ids]
    out_offsets = torch.zeros(ids.numel() + 1, dtype=torch.int64, device=chars.device)
    out_offsets[1:] = torch.cumsum(mine, 0)
    total = int(out_offsets[-1].item())
    out_chars = torch.empty(max(total, 1), dtype=torch.uint8, device=chars.device)
    block = 1 << 18
    for a in range(0, int(ids.numel()), block):
        b = min(int(ids.numel()), a + block)
        c0, c1 = int(out_offsets[a].item()), int(out_offsets[b].item())
        shift = torch.repeat_interleave(offsets[ids[a:b]] - out_offsets[a:b], mine[a:b])
        out_chars[c0:c1] = chars[shift + torch.arange(c0, c1, dtype=torch.int64, device=chars.device)]
    return out_chars[:total], out_offsets, ids
