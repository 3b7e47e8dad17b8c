#!/usr/bin/env python
"""Times the device builder for linear references (gcsa_b200_build_linear) and the creation of the device index
from its arrays, at one or more reference lengths.  GCSA_B200_VERBOSE=1 prints the builder's stages.

  python scripts/bench_build.py --mbp 100,1000,3000 [--check]      (--check: compare with the host builder, small sizes only)
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
    ap.add_argument("--mbp", default="100")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--kmer-table-k", type=int, default=16)
    ap.add_argument("--queries", type=int, default=10_000_000)
    ap.add_argument("--options", default="{}", help="JSON: several option sets for GCSA(...) as a list of objects, tried one after the other on each index")
    args = ap.parse_args()
    import torch
    from gcsa2_b200 import GCSA, synth
    from gcsa2_b200.builder import build_index, build_linear
    for mbp in [float(x) for x in args.mbp.split(",")]:
        L = int(mbp * 1_000_000)
        out = {"mbp": mbp}
        t0 = time.time()
        seq = synth.device_sequence(L, seed=4)
        torch.cuda.synchronize()
        out["sequence_s"] = time.time() - t0
        t0 = time.time()
        built = build_linear(seq, k=16, doubling_steps=3, raw=True)
        out["build_linear_s"] = time.time() - t0
        out["path_nodes"] = built.path_nodes
        if args.check:
            flat, lcp = built.flat(), built.lcp()
            host, hlcp, _ = build_index(synth.linear_graph(seq.cpu().numpy(), node_len=32), 16, 3)
            same = all((np.asarray(a) == np.asarray(b)).all() for a, b in
                       [(flat.C, host.C), (flat.edges, host.edges), (flat.sampled_paths, host.sampled_paths),
                        (flat.stored_samples, host.stored_samples), (flat.samples, host.samples), (lcp.data, hlcp.data)] +
                       [(flat.bwt[c], host.bwt[c]) for c in range(7)])
            out["same_as_host_builder"] = bool(same)
        n, length = args.queries, 32
        chars = synth.device_patterns(seq, n, length, seed=17)
        d_sp = torch.empty(n, dtype=torch.int64, device="cuda"); d_ep = torch.empty_like(d_sp)
        stream = torch.cuda.current_stream()
        option_sets = json.loads(args.options)
        for options in (option_sets if isinstance(option_sets, list) else [option_sets]):
            res = dict(out); res["options"] = options
            try:
                t0 = time.time()
                index = GCSA(built, device=0, kmer_table_k=args.kmer_table_k, **options)
                torch.cuda.synchronize()
                res["index_create_s"] = time.time() - t0
                res["device_bytes"] = index.deviceBytes(); res["fused_table"] = index.fusedTable(); res["two_step"] = index.twoStep(); res["jump_k"] = index.jumpK()
                for _ in range(3):
                    index.find_fixed_device(chars, length, n, d_sp, d_ep, stream.cuda_stream)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(5):
                    index.find_fixed_device(chars, length, n, d_sp, d_ep, stream.cuda_stream)
                e1.record(stream); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                res["find_ms"] = ms; res["find_gqps"] = n / ms / 1e6
                res["found"] = int(((d_sp + 1) <= (d_ep + 1)).sum().item())
                m = min(n, 1_000_000)
                _, _, st = index.find_fixed_batch(chars[:m * length].cpu().numpy(), length, stats=True)
                res["probes_per_query"] = (st["sector_probes"] + st["table_hits"]) / m; res["lf_steps_per_query"] = st["lf_steps"] / m
                index.close()
                del index
            except Exception as exc:
                res["error"] = "%s: %s" % (type(exc).__name__, exc)
            torch.cuda.empty_cache()
            print(json.dumps(res), flush=True)
        built.free()
        del chars, d_sp, d_ep, seq
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
