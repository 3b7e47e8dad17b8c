#!/usr/bin/env python
"""Differential fuzzing of the CUDA engine's source on the host emulation (tests/emu) against the CPU oracle.

Test infrastructure: random small graphs (linear references, SNP graphs with repeats, graphs with indels over two-
and four-letter alphabets), random creation options (k-mer table size, fused entries, jump tables, two-step blocks,
locate tables), random emulated SM counts, and patterns of every kind (walks through the graph, with substitutions,
with N / $ / lower case, uniform random, empty, short); find, count, locate (sorted, raw, bounded), parent, depth and
the MEM-style scan (with and without the jump variant) must equal the oracle bit for bit.

  python scripts/fuzz_emu.py --minutes 30 [--seed 1]
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()

    from emu import build_emu
    from gcsa2_b200 import capi
    capi._lib = capi._bind(ctypes.CDLL(build_emu.build()))
    from brute import random_graph
    from test_builder import flat_equal
    from gcsa2_b200 import GCSA, LCPArray, mem_batch, synth
    from gcsa2_b200.builder import CharGraph, build_index
    from oracle import oracle as orc

    deadline = time.time() + 60.0 * args.minutes
    rounds, seed = 0, args.seed
    while time.time() < deadline:
        rng = np.random.default_rng(seed)
        kind = int(rng.integers(0, 3))
        os.environ["GCSA_EMU_SMS"] = str(int(rng.integers(1, 4)))
        os.environ["GCSA_B200_FIND_REFILL"] = str(int(rng.choice([1, 8, 16, 24, 32])))
        os.environ["GCSA_B200_MEM_JUMP"] = str(int(rng.integers(0, 2)))
        os.environ["GCSA_B200_LOCATE_SMALL"] = str(int(rng.integers(0, 4) > 0))
        if kind == 0:                                                 # linear reference, maybe with repeats
            L = int(rng.integers(2_000, 60_000))
            seq = synth.random_sequence(L, seed=seed)
            for _ in range(int(rng.integers(0, 4))):
                a, b, w = int(rng.integers(0, L - 600)), int(rng.integers(0, L - 600)), int(rng.integers(20, 500))
                seq[a:a + w] = seq[b:b + w]
            graph = synth.linear_graph(seq)
            steps = int(rng.integers(1, 4))
            flat, flcp, _ = build_index(graph, 16, steps)
            if rounds % 4 == 0:                                       # the device builder for linear references emits the same arrays
                from gcsa2_b200.builder import build_linear
                dflat, dlcp = build_linear(seq, k=16, doubling_steps=steps)
                assert flat_equal(flat, dflat) == [] and (dlcp.data == flcp.data).all(), ("build_linear", seed, L, steps)
            sampler = lambda n, ln, s: synth.patterns_from_sequence(seq, n, ln, seed=s)
            what = "linear L=%d" % L
        elif kind == 1:                                               # SNP graph
            L = int(rng.integers(2_000, 40_000))
            seq = synth.random_sequence(L, seed=seed)
            for _ in range(int(rng.integers(0, 3))):
                a, b, w = int(rng.integers(0, L - 600)), int(rng.integers(0, L - 600)), int(rng.integers(20, 500))
                seq[a:a + w] = seq[b:b + w]
            rate = float(rng.choice([0.005, 0.01, 0.03, 0.08]))
            graph, sites, alt = synth.snp_graph(seq, seed=seed, snp_rate=rate)
            flat, flcp, _ = build_index(graph, 16, int(rng.integers(2, 4)))
            sampler = lambda n, ln, s: synth.patterns_from_snp_graph(seq, sites, alt, n, ln, seed=s)
            what = "snp L=%d rate=%g" % (L, rate)
        else:                                                         # small graph with indels, short kmers
            g = random_graph(rng, int(rng.integers(200, 2500)), 8, snp_rate=0.1, node_len=4,
                             alphabet=[(1, 2, 3, 4), (1, 2)][int(rng.integers(0, 2))])
            cg = CharGraph.from_lists(g.comps, g.values, g.succ, g.sources, g.sink)
            flat, flcp, _ = build_index(cg, 2, int(rng.integers(1, 4)), sample_period=int(rng.choice([4, 16, 64])),
                                        lcp_branching=int(rng.choice([2, 3, 4, 64])))
            sampler = None
            what = "indel graph N=%d" % flat.path_nodes
        N = flat.path_nodes
        opts = dict(kmer_table_k=int(rng.choice([0, 1, 2, 4, 6, 8, 9])), two_step=bool(rng.integers(0, 2)),
                    walk_table=[None, 0, 1, 2][int(rng.integers(0, 4))], jump_table=[False, True, True, "wide"][int(rng.integers(0, 4))],
                    fused_table=[None, True, False][int(rng.integers(0, 3))])
        gpu, ora = GCSA(flat, **opts), orc.OracleGCSA(flat)
        glcp, olcp = LCPArray(flcp), orc.OracleLCP(flcp)
        tag = (seed, what, opts, {k: os.environ[k] for k in ("GCSA_EMU_SMS", "GCSA_B200_FIND_REFILL", "GCSA_B200_MEM_JUMP", "GCSA_B200_LOCATE_SMALL")})

        # patterns
        pats = []
        alphabet = np.frombuffer(b"ACGTACGTACGTACGTacgtN$#x", dtype=np.uint8)
        if sampler is not None:
            for ln in rng.integers(1, 140, size=6):
                n = int(rng.integers(50, 800))
                c, o = sampler(n, int(ln), int(rng.integers(0, 1 << 30)))
                c = c.copy()
                for i in range(n):
                    r = rng.random()
                    if r < 0.3:
                        c[int(o[i]) + int(rng.integers(0, ln))] = alphabet[int(rng.integers(0, alphabet.size))]
                    elif r < 0.4:
                        c[int(o[i]):int(o[i + 1])] |= 0x20
                    pats.append(bytes(c[int(o[i]):int(o[i + 1])]))
        lengths = rng.integers(0, 40, size=(600 if rng.random() < 0.5 else 4200))     # (4096 patterns and more: the batch starts in the chain kernel)
        pats += [bytes(alphabet[rng.integers(0, alphabet.size if rng.random() < 0.3 else 16, size=int(ln))]) for ln in lengths]
        chars, offsets = orc.pack_patterns(pats)

        sp, ep = gpu.find_batch(chars, offsets)
        osp, oep, _ = ora.find_batch(chars, offsets, threads=4)
        assert (sp == osp).all() and (ep == oep).all(), ("find", tag, np.flatnonzero((sp != osp) | (ep != oep))[:5])

        # a batch of k-mers (one length, at least 4096 of them): the two-kernel form with its work lists
        if sampler is not None and opts["kmer_table_k"] > 0:
            os.environ["GCSA_B200_FIND_UNROLL"] = str(int(rng.choice([1, 2, 4])))
            ln = int(rng.choice([opts["kmer_table_k"], 12, 16, 24, 31, 32, 33, 50, 64, 101]))
            if ln >= opts["kmer_table_k"]:
                kc, ko = sampler(4500, ln, int(rng.integers(0, 1 << 30)))
                kc = kc.copy()
                for i in range(0, 4500, 3):
                    r = rng.random()
                    if r < 0.5:
                        kc[int(ko[i]) + int(rng.integers(0, ln))] = alphabet[int(rng.integers(0, alphabet.size if r < 0.1 else 16))]
                    elif r < 0.6:
                        kc[int(ko[i]):int(ko[i + 1])] = synth.random_patterns(1, ln, seed=int(rng.integers(0, 1 << 30)))[0]
                ksp0, kep0, _ = ora.find_batch(kc, ko, threads=4)
                ksp, kep = gpu.find_fixed_batch(kc, ln)
                assert (ksp == ksp0).all() and (kep == kep0).all(), ("k-mer form", ln, tag, np.flatnonzero((ksp != ksp0) | (kep != kep0))[:5])

        # fixed-length batches through the host entry point, with and without host-side 2-bit packing
        if sampler is not None and rounds % 8 == 0:
            ln = int(rng.choice([20, 32, 33, 45, 64]))
            n_fixed = 540_000 if rounds % 16 == 0 else 280_000        # three to five chunks: raw copies and packing share the batch
            fc, fo = sampler(n_fixed, ln, int(rng.integers(0, 1 << 30)))
            fc = fc.copy()
            rc, _ = synth.random_patterns(10_000, ln, seed=seed)
            fc[: rc.size] = rc
            if rng.random() < 0.5:
                fc[int(rng.integers(0, fc.size))] = ord("N")
            fsp0, fep0, _ = ora.find_batch(fc, fo, threads=4)
            for packing in ("0", "2"):
                os.environ["GCSA_B200_HOST_PACK"] = packing
                fsp, fep = gpu.find_fixed_batch(fc, ln)
                assert (fsp == fsp0).all() and (fep == fep0).all(), ("find_fixed", packing, tag)
            os.environ.pop("GCSA_B200_HOST_PACK")

        # ranges: what find() produced plus arbitrary ones
        a = rng.integers(0, N, size=500).astype(np.uint64)
        b = np.minimum(a + rng.integers(0, 20, size=500).astype(np.uint64), np.uint64(N + 2))
        rsp = np.concatenate([sp, a, np.array([3, 0], dtype=np.uint64)]); rep = np.concatenate([ep, b, np.array([2, N - 1], dtype=np.uint64)])
        assert (gpu.count_batch(rsp, rep) == np.array([ora.count((int(x), int(y))) for x, y in zip(rsp, rep)], dtype=np.uint64)).all(), ("count", tag)
        offs, vals = gpu.locate_batch(rsp, rep)
        ooffs, ovals, _ = ora.locate_batch(rsp, rep, threads=4)
        assert (offs == ooffs).all() and (vals == ovals).all(), ("locate", tag)
        roffs, rvals = gpu.locate_batch(rsp[:300], rep[:300], sort=False)
        for i in range(300):                                          # sort = false: the same positions before sort + unique
            raw = rvals[int(roffs[i]):int(roffs[i + 1])]
            assert sorted(set(int(x) for x in raw)) == [int(x) for x in vals[int(offs[i]):int(offs[i + 1])]], ("locate raw", tag, i)
        for i in rng.integers(0, rsp.size, size=20):
            r = (int(rsp[i]), int(rep[i]))
            assert list(gpu.locate(r, max_positions=5)) == list(ora.locate(r, max_positions=5)), ("locate max", tag, r)
        valid = (rep < N) & (rsp <= rep)
        par = glcp.parent_batch(rsp[valid], rep[valid])
        opar, _ = olcp.parent_batch(rsp[valid], rep[valid], threads=4)
        assert (par == opar).all(), ("parent", tag)
        dep = glcp.depth_batch(par[:200, 0].copy(), par[:200, 1].copy())
        assert list(dep) == [olcp.depth((int(x), int(y))) for x, y in zip(par[:200, 0], par[:200, 1])], ("depth", tag)

        moffs, mvals = mem_batch(gpu, glcp, chars, offsets)
        eoffs, evals, _ = orc.mem_batch(ora, olcp, chars, offsets, threads=4)
        assert (moffs == eoffs).all() and mvals.shape == evals.shape and (mvals == evals).all(), ("mem", tag)

        gpu.close(); glcp.close()
        rounds += 1; seed += 1
        if rounds % 10 == 0:
            print("%d rounds ok (last: %s)" % (rounds, what), flush=True)
    print("fuzz_emu: %d rounds, no difference" % rounds)


if __name__ == "__main__":
    main()
