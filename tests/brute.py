"""Brute-force GCSA construction for tiny graphs (test infrastructure).

Builds the index straight from the definitions, with explicit strings and no cleverness:
every node's order-K path labels are enumerated, sorted, pruned to the maximal subtrees
whose strings share one value set (paper/paper.tex:246-254; src/path_graph.cpp:577-609,
1154-1226), and the arrays are emitted by the rules of src/gcsa.cpp:568-704.
It validates the scalable builder (gcsa2_b200.builder) and produces fixtures whose
contents can be checked by eye.
"""
import bisect
from dataclasses import dataclass

import numpy as np

from gcsa2_b200.flat import FlatGCSA, FlatLCP, SIGMA, bits_from_positions

SINK_COMP = 0
SOURCE_COMP = 6
U64 = (1 << 64) - 1


@dataclass
class SimpleGraph:
    """Single-character nodes.  sink's label is '$' (comp 0); the technical edge sink -> source
    (paper/paper.tex:179) is implied: every node in `sources` gets predecessor '$'."""
    comps: list        # comp value per node
    values: list       # node_type per node
    succ: list         # list of successor lists
    sources: list      # nodes preceded by the sink through the technical edge
    sink: int


def kstrings(graph, v, K):
    """All order-K labels of paths starting at v; past the sink the label is padded with '$'."""
    out = set()
    stack = [(v, (graph.comps[v],))]
    while stack:
        node, s = stack.pop()
        if len(s) == K:
            out.add(s); continue
        if node == graph.sink:
            out.add(s + (SINK_COMP,) * (K - len(s))); continue
        for w in graph.succ[node]:
            stack.append((w, s + (graph.comps[w],)))
    return out


def lcp_of(a, b):
    n = 0
    while n < len(a) and n < len(b) and a[n] == b[n]:
        n += 1
    return n


def redundant_counts(node_values, lcp):
    """R[] of Sadakane's counting structure by the in-order suffix-tree walk of
    src/gcsa.cpp:590-619 (stack of (lcp, first_time, last_time); LCP[0] handled as -1)."""
    N = len(node_values)
    red = [0] * max(0, N - 1)
    prev_occ, node_lcp, first_time, last_time = {}, [], [], []
    for i in range(N):
        curr_lcp = lcp[i] + (1 if i > 0 else 0)
        while node_lcp and node_lcp[-1] > curr_lcp:
            node_lcp.pop(); first_time.pop(); last_time.pop()
        if node_lcp and node_lcp[-1] == curr_lcp:
            last_time[-1] = i
        else:
            node_lcp.append(curr_lcp); first_time.append(i); last_time.append(i)
        for x in node_values[i]:
            p = prev_occ.get(x, 0)
            if p > 0:
                pos = bisect.bisect_left(last_time, p)
                red[first_time[pos] - 1] += 1
            prev_occ[x] = i + 1
    return red


class BruteIndex:
    def __init__(self, graph, K, sample_period=64):
        self.graph, self.K = graph, K
        n = len(graph.comps)
        preds = [[] for _ in range(n)]
        for u in range(n):
            if u == graph.sink:
                continue
            for w in graph.succ[u]:
                preds[w].append(u)
        pred_comps = [set(graph.comps[u] for u in preds[w]) for w in range(n)]
        for s in graph.sources:
            pred_comps[s].add(SINK_COMP)

        table = {}
        for v in range(n):
            for s in kstrings(graph, v, K):
                table.setdefault(s, set()).add(v)
        strings = sorted(table)
        vsets = [frozenset(table[s]) for s in strings]
        self.strings = strings

        # maximal pruning: trie descent, stop where all strings share one value set
        nodes = []   # (lo, hi, key)
        def prune(lo, hi, depth):
            if all(vsets[i] == vsets[lo] for i in range(lo, hi)):
                nodes.append((lo, hi, strings[lo][:depth]))
                return
            i = lo
            while i < hi:
                j = i
                while j < hi and strings[j][depth] == strings[i][depth]:
                    j += 1
                prune(i, j, depth + 1)
                i = j
        prune(0, len(strings), 0)
        self.nodes = nodes
        N = len(nodes)
        starts = [lo for lo, _, _ in nodes]
        def node_of_string(s):
            idx = bisect.bisect_left(strings, s)
            assert idx < len(strings) and strings[idx] == s, "predecessor string missing"
            return bisect.bisect_right(starts, idx) - 1

        self.keys = [key for _, _, key in nodes]
        node_vsets = [vsets[lo] for lo, _, _ in nodes]
        self.node_values = [sorted(set(graph.values[v] for v in vs)) for vs in node_vsets]

        # BWT bits and edges (src/gcsa.cpp:568-588, 689-696)
        bwt = [[] for _ in range(SIGMA)]
        outdeg = [0] * N
        pred_node = [dict() for _ in range(N)]
        for i, (lo, hi, _) in enumerate(nodes):
            cs = set()
            for w in node_vsets[i]:
                cs |= pred_comps[w]
            for c in sorted(cs):
                bwt[c].append(i)
        self.consistent = True
        for c in range(SIGMA):
            last_j = -1
            for i in bwt[c]:
                lo, hi, _ = nodes[i]
                js = set()
                for idx in range(lo, hi):
                    s = (c,) + strings[idx][:K - 1]
                    if c == SINK_COMP:
                        s = (SINK_COMP,) * K       # the sink's own label is "$$$..." by convention
                    js.add(node_of_string(s))
                if len(js) != 1:
                    self.consistent = False
                j = min(js)
                if j < last_j:
                    self.consistent = False
                last_j = j
                pred_node[i][c] = j
                outdeg[j] += 1
        if any(d == 0 for d in outdeg):
            self.consistent = False
        C = [0] * (SIGMA + 1)
        for c in range(SIGMA):
            C[c + 1] = C[c] + len(bwt[c])
        edge_ones, total = [], 0
        for i in range(N):
            total += outdeg[i]
            edge_ones.append(total - 1)
        self.outdeg = outdeg

        # LCP array between adjacent nodes (src/path_graph.cpp:1204)
        lcp = [0] * N
        for i in range(1, N):
            lcp[i] = lcp_of(strings[nodes[i][0] - 1], strings[nodes[i][0]])
        self.lcp = lcp

        # samples (src/gcsa.cpp:621-658)
        sampled, stored, last_bits = [], [], []
        for i in range(N):
            cur = self.node_values[i]
            cs = sorted(pred_node[i].keys())
            sample = len(cs) > 1 or SINK_COMP in cs or any(v % sample_period == 0 for v in cur)
            if not sample:
                pv = self.node_values[pred_node[i][cs[0]]]
                if len(pv) != len(cur) or any(cur[k] != (pv[k] + 1) & U64 for k in range(len(cur))):
                    sample = True
            if sample:
                sampled.append(i)
                stored.extend(cur)
                last_bits.append(len(stored) - 1)

        # counting structures (src/gcsa.cpp:590-619, 671-672)
        occ = [len(v) - 1 for v in self.node_values]
        red = redundant_counts(self.node_values, lcp)
        self.occ, self.red = occ, red

        filt = [i for i in range(N) if occ[i] > 0]
        ev_ones, tail = [], 0
        for i in filt:
            tail += occ[i]; ev_ones.append(tail - 1)
        rd_ones, tail = [], 0
        for i in range(N - 1):
            tail += red[i] + 1; rd_ones.append(tail - 1)

        self.flat = FlatGCSA(
            path_nodes=N, edge_count=total, order=K, C=np.array(C, dtype=np.uint64),
            bwt=[bits_from_positions(bwt[c], N) for c in range(SIGMA)],
            edges=bits_from_positions(edge_ones, total),
            sampled_paths=bits_from_positions(sampled, N),
            sample_count=len(stored), stored_samples=np.array(stored, dtype=np.uint64),
            samples=bits_from_positions(last_bits, len(stored)),
            extra_filter=bits_from_positions(filt, N),
            extra_values_len=sum(occ), extra_values=bits_from_positions(ev_ones, sum(occ)),
            redundant_len=(N - 1) + sum(red), redundant=bits_from_positions(rd_ones, (N - 1) + sum(red)))
        self.flat_lcp = FlatLCP.from_values(np.array(lcp, dtype=np.uint8), branching=4)
        self.bwt_sets = bwt
        self.sampled = sampled

    def find(self, pattern_comps):
        """Definition of find(): nodes whose key has the pattern as a prefix or vice versa, as a
        closed range -- valid for patterns no longer than K.  Empty -> None."""
        p = tuple(pattern_comps)
        hits = [i for i, (lo, hi, _) in enumerate(self.nodes)
                if any(self.strings[idx][:len(p)] == p for idx in range(lo, hi))]
        if not hits:
            return None
        assert hits == list(range(hits[0], hits[-1] + 1))
        return (hits[0], hits[-1])


def random_graph(rng, length, K, snp_rate=0.15, ins_rate=0.05, node_len=4, alphabet=(1, 2, 3, 4)):
    """A backbone with SNP bubbles (and a few extra alternative branches), vg-like node ids:
    one id per `node_len` backbone characters, alt alleles get their own ids."""
    from gcsa2_b200.flat import node_encode
    comps, values, succ = [], [], []
    def add(comp, value):
        comps.append(comp); values.append(value); succ.append([])
        return len(comps) - 1
    next_id = [1]
    src = add(SOURCE_COMP, node_encode(next_id[0], 0)); next_id[0] += 1
    prev_ends = [src]
    cur_id, cur_off = next_id[0], 0; next_id[0] += 1
    for _ in range(length):
        c = int(rng.choice(alphabet))
        if cur_off >= node_len or len(prev_ends) > 1:
            cur_id, cur_off = next_id[0], 0; next_id[0] += 1
        ref = add(c, node_encode(cur_id, cur_off)); cur_off += 1
        for p in prev_ends:
            succ[p].append(ref)
        ends = [ref]
        if rng.random() < snp_rate:
            alt_c = int(rng.choice([a for a in alphabet if a != c]))
            alt = add(alt_c, node_encode(next_id[0], 0)); next_id[0] += 1
            for p in prev_ends:
                succ[p].append(alt)
            ends.append(alt)
            cur_off = node_len   # force a new id after the bubble
        prev_ends = ends
    sink = add(SINK_COMP, node_encode(next_id[0], 0))
    for p in prev_ends:
        succ[p].append(sink)
    return SimpleGraph(comps=comps, values=values, succ=succ, sources=[src], sink=sink)
