#!/usr/bin/env python
"""find() of 64-mers on the configs[2] SNP graph: what the kernels execute (steps, probes, work lists) and how long they
take.  Run under ncu for the per-kernel times:  ncu --metrics gpu__time_duration.sum -k regex:find_ python scripts/probe_graph_find.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    from gcsa2_b200 import GCSA, synth
    from gcsa2_b200.builder import build_index
    mbp = float(os.environ.get("MBP", "50")); n = int(os.environ.get("QUERIES", "4000000")); length = 64
    seq = synth.random_sequence(int(mbp * 1e6), seed=3)
    graph, sites, alt = synth.snp_graph(seq, seed=3, snp_rate=0.01)
    flat, _, _ = build_index(graph, 16, 3)
    chars = np.empty(n * length, dtype=np.uint8)
    for i, q0 in enumerate(range(0, n, 1_000_000)):
        m = min(1_000_000, n - q0)
        c, _ = synth.patterns_from_snp_graph(seq, sites, alt, m, length, seed=700 + i)
        chars[q0 * length:(q0 + m) * length] = c
    for k in [int(x) for x in os.environ.get("TABLE_K", "12,14").split(",")]:
        index = GCSA(flat, kmer_table_k=k)
        if os.environ.get("STATS", "1") != "0":
            sp, ep, stats = index.find_fixed_batch(chars, length, stats=True)
            print("kmer_table_k", k, "found", int((sp <= ep).sum()), "stats", stats, flush=True)
        else:
            sp, ep = index.find_fixed_batch(chars, length)
        d_chars = torch.from_numpy(chars).cuda()
        d_sp = torch.empty(n, dtype=torch.int64, device="cuda"); d_ep = torch.empty_like(d_sp)
        stream = torch.cuda.current_stream()
        for _ in range(3):
            index.find_fixed_device(d_chars, length, n, d_sp, d_ep, stream.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            index.find_fixed_device(d_chars, length, n, d_sp, d_ep, stream.cuda_stream)
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("kmer_table_k", k, "ms", ms, "queries/s", n / ms * 1e3, "range length mean", float((ep - sp + 1).mean()), flush=True)
        # the same patterns through the general kernel alone (the offsets form)
        d_off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * length
        for _ in range(3):
            index.find_device(d_chars, d_off, n, d_sp, d_ep, stream.cuda_stream)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(5):
            index.find_device(d_chars, d_off, n, d_sp, d_ep, stream.cuda_stream)
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("kmer_table_k", k, "general kernel alone: ms", ms, "queries/s", n / ms * 1e3, flush=True)
        # mixed lengths 16..256 (the configs[4] patterns), as generated and sorted by length
        mc, mo = synth.mixed_length_patterns(seq, sites, alt, 1_000_000, 16, 256, seed=5, error_rate=0.0)
        lens = np.diff(mo.astype(np.int64))
        order = np.argsort(lens, kind="stable")
        sc = np.concatenate([mc[int(mo[i]):int(mo[i + 1])] for i in order]) if os.environ.get("SORTED", "1") != "0" else None
        for tag, cc, oo in (("mixed 16..256", mc, mo), ("mixed, sorted by length", sc, np.concatenate([[0], np.cumsum(lens[order])]).astype(np.uint64))):
            if cc is None:
                continue
            m = oo.size - 1
            dc = torch.from_numpy(cc).cuda(); do = torch.from_numpy(oo.view(np.int64)).cuda()
            ds = torch.empty(m, dtype=torch.int64, device="cuda"); de = torch.empty_like(ds)
            for _ in range(3):
                index.find_device(dc, do, m, ds, de, stream.cuda_stream)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(5):
                index.find_device(dc, do, m, ds, de, stream.cuda_stream)
            e1.record(stream); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print("kmer_table_k", k, tag, "ms", ms, "queries/s", m / ms * 1e3, "found", int((ds <= de).sum().item()), flush=True)
        index.close()


if __name__ == "__main__":
    main()
