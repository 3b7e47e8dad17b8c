"""The reference's self-consistency check, verifyIndex (src/algorithms.cpp:101-295), restated
over any object with the reference's query interface (the oracle or the CUDA engine).

For every distinct kmer label of the construction input:
  1. find(kmer) is non-empty (kmer truncated after its first '$', :127-129, :132-143);
  2. parent(range) equals the first different range obtained by dropping characters from the
     right end of the kmer and parent.lcp() equals that shorter length; depth(parent.range())
     equals parent.lcp() (:146-181);
  3. count(range) = number of distinct start positions of the label (:183-200);
  4. locate(range) = that sorted distinct set (:202-234);
  5. locate(range, 10) has min(10, n) values and is a subset of it (:236-274).
"""
import numpy as np

RANDOM_LOCATE_SIZE = 10          # src/algorithms.cpp:60
COMP2CHAR = "$ACGTN#"


def kmer_table(kmers):
    """-> list of (pattern string, sorted distinct from values)"""
    labels = (kmers.key >> np.uint64(16))
    order = np.argsort(labels, kind="stable")
    labels, frm = labels[order], kmers.from_[order]
    out = []
    i = 0
    n = labels.size
    bounds = np.flatnonzero(np.concatenate([[True], labels[1:] != labels[:-1], [True]]))
    for a, b in zip(bounds[:-1], bounds[1:]):
        x = int(labels[a])
        s = "".join(COMP2CHAR[(x >> (3 * (kmers.k - 1 - i))) & 7] for i in range(kmers.k))
        p = s.find("$")
        if p >= 0:
            s = s[:p + 1]
        out.append((s, sorted(set(int(v) for v in frm[a:b]))))
    return out


def is_empty(rng):
    m = (1 << 64) - 1
    return ((rng[0] + 1) & m) > ((rng[1] + 1) & m)


def verify_index(index, lcp, table, limit=None, seed=0):
    """Returns a list of failure descriptions (empty = index verification complete)."""
    fails = []
    if limit is not None and len(table) > limit:
        pick = np.random.default_rng(seed).choice(len(table), size=limit, replace=False)
        table = [table[i] for i in sorted(pick)]
    for kmer, expected in table:
        rng = index.find(kmer)
        if is_empty(rng):
            fails.append("find(%s) returned empty range" % kmer); continue
        if lcp is not None:
            parent = lcp.parent(rng)
            query, end = rng, len(kmer)
            while query == rng:
                end -= 1
                query = index.find(kmer[:end])
            if (parent[0], parent[1]) != query or parent[4] != end:
                fails.append("parent%s returned %s, expected %s at depth %d" % (rng, parent, query, end)); continue
            depth = lcp.depth((parent[0], parent[1]))
            if depth != parent[4]:
                fails.append("depth%s returned %d, expected %d" % ((parent[0], parent[1]), depth, parent[4])); continue
        c = index.count(rng)
        if c != len(expected):
            fails.append("count%s: expected %d, got %d (%s)" % (rng, len(expected), c, kmer)); continue
        occs = index.locate(rng)
        if occs != expected:
            fails.append("locate(%s): expected %s, got %s" % (kmer, expected[:5], occs[:5])); continue
        rnd = index.locate(rng, RANDOM_LOCATE_SIZE)
        if len(rnd) != min(RANDOM_LOCATE_SIZE, len(occs)) or not set(rnd) <= set(occs) or rnd != sorted(rnd):
            fails.append("locate(%s, %d): %s is not a sorted subset of size %d" % (kmer, RANDOM_LOCATE_SIZE, rnd, min(RANDOM_LOCATE_SIZE, len(occs))))
    return fails
