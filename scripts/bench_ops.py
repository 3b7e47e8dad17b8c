#!/usr/bin/env python
"""Per-operation measurements on BASELINE.json configs[2] (SNP-bubble variation graph, order 128):
find(), count(), locate(), parent(), depth() on the GPU (device-resident, CUDA events) next to the
CPU oracle on the host cores.  Writes one JSON object per operation.

  python scripts/bench_ops.py [--mbp 50] [--queries 10000000] [--length 64] [--out gpurun_out/ops.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=50.0)
    ap.add_argument("--queries", type=int, default=10_000_000)
    ap.add_argument("--length", type=int, default=64)
    ap.add_argument("--snp-rate", type=float, default=0.01)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    ap.add_argument("--kmer-table-k", type=int, default=14)
    ap.add_argument("--out", default="")
    ap.add_argument("--ops", default="count,locate,parent,mem,kmers,compare,verify",
                    help="comma-separated subset of: count,locate,locate_short,parent,mem,kmers,compare,verify (find always runs)")
    args = ap.parse_args()

    import torch
    from gcsa2_b200 import GCSA, LCPArray, synth
    from gcsa2_b200.builder import build_index
    from oracle import oracle as orc

    L, n, length = int(args.mbp * 1e6), args.queries, args.length
    t0 = time.time()
    seq = synth.random_sequence(L, seed=3)
    graph, sites, alt = synth.snp_graph(seq, seed=3, snp_rate=args.snp_rate)
    flat, flcp, kmers = build_index(graph, 16, 3)
    build_s = time.time() - t0
    ops = set(args.ops.split(","))
    chars = np.empty(n * length, dtype=np.uint8)
    for i, q0 in enumerate(range(0, n, 1_000_000)):
        m = min(1_000_000, n - q0)
        c, _ = synth.patterns_from_snp_graph(seq, sites, alt, m, length, seed=700 + i)
        chars[q0 * length:(q0 + m) * length] = c
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)

    index = GCSA(flat, kmer_table_k=args.kmer_table_k)
    lcp = LCPArray(flcp)
    ora, olcp = orc.OracleGCSA(flat), orc.OracleLCP(flcp)
    threads = orc.lib().oracle_max_threads()
    stream = torch.cuda.current_stream()
    results = []

    def timed(fn, steps=args.steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def report(op, unit, units_per_step, ms, cpu_value, cpu_sample, parity, extra=None):
        row = {"op": op, "unit": unit, "gpu_value": units_per_step / (ms / 1000.0), "gpu_ms_per_step": ms,
               "cpu_value": cpu_value, "cpu_cores": threads, "cpu_sample": cpu_sample, "speedup": units_per_step / (ms / 1000.0) / cpu_value,
               "parity_on_sample": bool(parity)}
        row.update(extra or {})
        results.append(row)
        print(json.dumps(row), flush=True)

    # ---- find ----
    d_chars = torch.from_numpy(chars).cuda()
    d_sp = torch.empty(n, dtype=torch.int64, device="cuda"); d_ep = torch.empty_like(d_sp)
    ms = timed(lambda: index.find_fixed_device(d_chars, length, n, d_sp, d_ep, stream.cuda_stream))
    sp = d_sp.cpu().numpy().view(np.uint64); ep = d_ep.cpu().numpy().view(np.uint64)
    m = min(n, args.cpu_sample)
    osp, oep, secs = ora.find_batch(chars[:m * length], offsets[:m + 1], threads=threads)
    report("find (%d-mers from walks through the graph)" % length, "queries/s", n, ms, m / secs, m,
           (osp == sp[:m]).all() and (oep == ep[:m]).all(),
           {"found": int(np.count_nonzero(sp <= ep)), "index": {"path_nodes": index.size(), "edges": index.edgeCount(),
                                                               "device_bytes": index.deviceBytes(), "two_step": index.twoStep(),
                                                               "kmer_table_k": index.kmerTableK()},
            "build_s": build_s})

    # ---- count ----  (locate needs the counts)
    d_cnt = torch.empty(n, dtype=torch.int64, device="cuda")
    ms = timed(lambda: index.count_device(d_sp, d_ep, n, d_cnt, stream.cuda_stream))
    cnt = d_cnt.cpu().numpy().view(np.uint64)
    ocnt, secs = ora.count_batch(sp[:m], ep[:m], threads=threads)
    report("count", "ranges/s", n, ms, m / secs, m, (ocnt == cnt[:m]).all())

    # ---- locate ----
    if "locate" in ops:
      total = int(cnt.sum())
      d_offs = torch.empty(n + 1, dtype=torch.int64, device="cuda")
      d_vals = torch.empty(total + 16, dtype=torch.int64, device="cuda")
      got = [0]
      def do_locate():
          got[0] = index.locate_device(d_sp, d_ep, n, d_offs, d_vals, total + 16, stream.cuda_stream)
      ms = timed(do_locate, steps=3)
      offs = d_offs.cpu().numpy().view(np.uint64); vals = d_vals[:got[0]].cpu().numpy().view(np.uint64)
      ml = min(m, 1_000_000)
      ooffs, ovals, secs = ora.locate_batch(sp[:ml], ep[:ml], threads=threads)
      k = int(ooffs[ml])
      report("locate (sorted distinct positions per range)", "positions/s", got[0], ms, k / secs, ml,
             (offs[:ml + 1] == ooffs).all() and (vals[:k] == ovals).all() and got[0] == total,
             {"positions": got[0], "ranges": n})

    # ---- locate() of short patterns: wide ranges, the general pipeline (segmented sort) instead of the short-range path ----
    if "locate_short" in ops:
      for plen, nq in ((12, 2_000_000), (10, 1_000_000), (8, 200_000), (7, 50_000), (6, 20_000)):
          pchars, poffsets = synth.patterns_from_snp_graph(seq, sites, alt, nq, plen, seed=600 + plen)
          psp, pep = index.find_batch(pchars, poffsets)
          d_psp = torch.from_numpy(psp.view(np.int64)).cuda(); d_pep = torch.from_numpy(pep.view(np.int64)).cuda()
          d_pcnt = torch.empty(nq, dtype=torch.int64, device="cuda")
          index.count_device(d_psp, d_pep, nq, d_pcnt, stream.cuda_stream); torch.cuda.synchronize()
          ptotal = int(d_pcnt.sum().item())
          d_poffs = torch.empty(nq + 1, dtype=torch.int64, device="cuda")
          d_pvals = torch.empty(ptotal + 16, dtype=torch.int64, device="cuda")
          pgot = [0]
          def do_locate_short():
              pgot[0] = index.locate_device(d_psp, d_pep, nq, d_poffs, d_pvals, ptotal + 16, stream.cuda_stream)
          ms = timed(do_locate_short, steps=3)
          ml = min(nq, 100_000)
          ooffs, ovals, secs = ora.locate_batch(psp[:ml], pep[:ml], threads=threads)
          kk = int(ooffs[ml])
          poffs = d_poffs.cpu().numpy().view(np.uint64); pvals = d_pvals[:kk].cpu().numpy().view(np.uint64)
          report("locate of %d-mers (%.1f path nodes per range)" % (plen, float((pep - psp + 1).astype(np.float64).mean())), "positions/s", pgot[0], ms, kk / secs, ml,
                 (poffs[:ml + 1] == ooffs).all() and (pvals == ovals).all() and pgot[0] == ptotal, {"positions": pgot[0], "ranges": nq})
          del d_psp, d_pep, d_pcnt, d_poffs, d_pvals

    # ---- parent / depth ----
    if "parent" in ops:
      d_par = torch.empty((n, 5), dtype=torch.int64, device="cuda")
      ms = timed(lambda: lcp.parent_device(d_sp, d_ep, n, d_par, stream.cuda_stream))
      par = d_par.cpu().numpy().view(np.uint64)
      opar, secs = olcp.parent_batch(sp[:m], ep[:m], threads=threads)
      report("parent", "ranges/s", n, ms, m / secs, m, (opar == par[:m]).all())
      d_dep = torch.empty(n, dtype=torch.int64, device="cuda")
      psp = torch.from_numpy(par[:, 0].copy().view(np.int64)).cuda(); pep = torch.from_numpy(par[:, 1].copy().view(np.int64)).cuda()
      ms = timed(lambda: lcp.depth_device(psp, pep, n, d_dep, stream.cuda_stream))
      dep = d_dep.cpu().numpy().view(np.uint64)
      md = min(m, 200_000)
      t0 = time.time()
      odep = np.array([olcp.depth((int(a), int(b))) for a, b in zip(par[:md, 0], par[:md, 1])], dtype=np.uint64)
      secs = time.time() - t0
      report("depth (of the parents)", "ranges/s", n, ms, md / secs, md, (odep == dep[:md]).all(), {"cpu_note": "single thread through ctypes"})

    # ---- MEM-style scan (config 5): mixed lengths 16..256, 1 % substitutions ----
    if "mem" in ops:
      from gcsa2_b200 import mem_device
      nm = min(n, 4_000_000)
      mchars, moffsets = synth.mixed_length_patterns(seq, sites, alt, nm, 16, 256, seed=900, error_rate=0.01)
      d_mchars = torch.from_numpy(mchars).cuda(); d_moff = torch.from_numpy(moffsets.view(np.int64)).cuda()
      d_moffs_out = torch.empty(nm + 1, dtype=torch.int64, device="cuda")
      cap = 16 * nm
      d_matches = torch.empty((cap, 4), dtype=torch.int64, device="cuda")
      got = [0]
      def do_mem():
          got[0] = mem_device(index, lcp, d_mchars, d_moff, nm, d_moffs_out, d_matches, cap, stream.cuda_stream)
      mm = min(nm, 400_000)
      eoffs, evals, secs = orc.mem_batch(ora, olcp, mchars[:int(moffsets[mm])], moffsets[:mm + 1], threads=threads)
      k = int(eoffs[mm])
      for jump in ("0", "1"):                                        # mem_kernel<.., JUMP>: singleton ranges follow the jump tables
          os.environ["GCSA_B200_MEM_JUMP"] = jump
          ms = timed(do_mem, steps=3)
          moffs = d_moffs_out.cpu().numpy().view(np.uint64); mvals = d_matches[:got[0]].cpu().numpy().view(np.uint64)
          report("MEM-style scan (LF + parent), lengths 16..256, GCSA_B200_MEM_JUMP=" + jump, "patterns/s", nm, ms, mm / secs, mm,
                 (moffs[:mm + 1] == eoffs).all() and (mvals[:k] == evals).all(),
                 {"matches": got[0], "pattern_bytes": int(moffsets[-1])})
      os.environ.pop("GCSA_B200_MEM_JUMP", None)

    # ---- countKMers ----
    for k in ((12, 16) if "kmers" in ops else ()):
        index.count_kmers(k); torch.cuda.synchronize()                              # warm-up: the stream-ordered pool grows to its working size
        t0 = time.time(); g = index.count_kmers(k); torch.cuda.synchronize(); gs = time.time() - t0
        t0 = time.time(); c = ora.count_kmers(k, threads=threads); cs = time.time() - t0
        row = {"op": "countKMers(k=%d)" % k, "unit": "kmers/s", "gpu_value": g / gs, "gpu_ms_per_step": gs * 1000.0,
               "cpu_value": c / cs, "cpu_cores": threads, "cpu_sample": "whole index", "speedup": cs / gs, "parity_on_sample": bool(g == c), "kmers": g}
        results.append(row); print(json.dumps(row), flush=True)

    # ---- compareKMers: the variation graph against its own backbone (src/algorithms.cpp:535-616) ----
    if "compare" in ops:
        bflat, _, _ = build_index(synth.linear_graph(seq, node_len=32), 16, 3)
        backbone = GCSA(bflat, kmer_table_k=0)
        obackbone = orc.OracleGCSA(bflat)
        for k in (12, 16):
            index.compare_kmers(backbone, k); torch.cuda.synchronize()                # warm-up at the same size (memory pool)
            t0 = time.time(); g = index.compare_kmers(backbone, k); torch.cuda.synchronize(); gs = time.time() - t0
            t0 = time.time(); c = ora.compare_kmers(obackbone, k, threads=threads)[0]; cs = time.time() - t0
            row = {"op": "compareKMers(k=%d): graph vs its backbone" % k, "unit": "kmers/s", "gpu_value": sum(g) / gs, "gpu_ms_per_step": gs * 1000.0,
                   "cpu_value": sum(c) / cs, "cpu_cores": threads, "cpu_sample": "both indexes whole", "speedup": cs / gs,
                   "parity_on_sample": bool(tuple(g) == tuple(c)), "shared_left_right": list(g)}
            results.append(row); print(json.dumps(row), flush=True)
        del backbone

    # ---- verifyIndex (src/algorithms.cpp:101-295), batched on the device ----
    if "verify" in ops:
        rep = index.verify(kmers, lcp)
        row = {"op": "verifyIndex (find, parent, depth, count, locate, locate(.,10) for every kmer label)", "unit": "patterns/s",
               "gpu_value": rep["unique"] / rep["seconds"], "gpu_ms_per_step": rep["seconds"] * 1000.0, "patterns": rep["unique"],
               "failures": rep["failures"], "kmer_records": int(kmers.key.size),
               "engine_seconds": rep["engine_seconds"],
               "note": "device-resident: records sorted, patterns built and five of the six predicates compared on the GPU; locate(range, 10) draws on the host"}
        from oracle import reference as ref
        if ref.available():
            # the reference's own verifyIndex needs its own (disk-based) construction: timed on a 1 Mbp graph
            small_graph, _, _ = synth.snp_graph(synth.random_sequence(1_000_000, seed=5), seed=5, snp_rate=args.snp_rate)
            sflat, slcp, skmers = build_index(small_graph, 16, 3)
            r = ref.ReferenceIndex.build(skmers, 3)
            t0 = time.time(); ok = r.verify(); cs = time.time() - t0
            srep = GCSA(sflat, kmer_table_k=8).verify(skmers, LCPArray(slcp))
            row.update({"cpu_value": srep["unique"] / cs, "cpu_cores": threads, "cpu_sample": "1 Mbp graph, %d labels, the reference's own verifyIndex (its sources + SDSL shim), ok=%s" % (srep["unique"], ok),
                        "gpu_value_same_sample": srep["unique"] / srep["seconds"], "parity_on_sample": bool(ok and srep["failures"] == 0)})
            row["speedup"] = row["gpu_value_same_sample"] / row["cpu_value"]
        results.append(row); print(json.dumps(row), flush=True)

    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
