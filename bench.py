#!/usr/bin/env python
"""bench.py -- k-mer find() throughput of the B200 engine on BASELINE.json configs[1]:
10 M 32-mers sampled from a 100 Mbp synthetic linear reference, order-128 index, per GPU.

  python bench.py --gpus N --steps K --warmup W            (torchrun launches N ranks for N > 1)
  python bench.py --impl reference ...                     (the CPU path on the host cores)

A "step" is one pass of find() over the rank's whole batch.  `value` is measured with the batch
resident in HBM (CUDA events on the launching stream); `e2e` through the C-ABI host entry point
gcsa_b200_find_host with pinned host buffers (H2D and D2H inside the timed region).
Weak scaling: every rank searches its own batch of the same size against its own replica of the
index; the only collective is the all-reduce of the result counters and of the step time (max).

Only the cpu_baseline leg and --impl reference execute oracle/ (as the thing timed on the CPU,
never on the product path).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1; the host-side index build (rank 0) and the CPU baseline are
# OpenMP code that should see the box's cores.  Must happen before the libraries are loaded.
if os.environ.get("OMP_NUM_THREADS") == "1" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

METRIC = "kmer_find_queries_per_sec"
UNIT = "queries/s"
# Dependent chains of random 32-byte sector reads over footprints beyond 4 GB: 43.4-44.7 G probes/s on a B200
# (scripts/gather_probe.cu, profiles/r01_random_probe_microbench.txt) -- the ceiling of a kernel whose work is random probes.
PROBE_CEILING = 44.0e9
PROBE_CEILING_SOURCE = "profiles/r01_random_probe_microbench.txt (scripts/gather_probe.cu, footprints of 4.8-19 GB)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-mbp", type=float, default=100.0, help="length of the synthetic linear reference (Mbp)")
    ap.add_argument("--queries", type=int, default=10_000_000, help="patterns per GPU")
    ap.add_argument("--pattern-length", type=int, default=32)
    ap.add_argument("--kmer-table-k", type=int, default=16, help="k-mer table: find() of all 4^k ACGT strings, 8 B each (16 = 34 GB; profiles/r01_kmer_table_sweep.txt)")
    ap.add_argument("--fused-table", type=int, default=-1, help="k-mer table entries of 16 B that carry the first jump (one load per 32-mer with k = 16; 69 GB): 1 = yes, 0 = no, -1 = the engine decides by free memory")
    ap.add_argument("--two-step", type=int, default=-1, help="1 = build and use the two-step blocks, 0 = never, -1 = by index size")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--index-cache", default="", help="npz file: load the built index from it if present, else save it there")
    ap.add_argument("--no-locate", action="store_true", help="skip the locate() leg (BASELINE.json configs[2])")
    ap.add_argument("--locate-mbp", type=float, default=50.0, help="backbone length of the SNP-bubble graph of the locate() leg (Mbp)")
    ap.add_argument("--locate-queries", type=int, default=10_000_000, help="64-mers per GPU in the locate() leg")
    ap.add_argument("--host-builder", action="store_true", help="build the cfg2 index with the host builder (builder.cpp, ~55 s) instead of the device builder")
    ap.add_argument("--no-wide-locate", action="store_true", help="skip locate() of the ranges of short patterns (part of the locate leg)")
    ap.add_argument("--no-mem", action="store_true", help="skip the configs[4] leg (MEM-style scan of mixed-length patterns)")
    ap.add_argument("--mem-patterns", type=int, default=4_000_000, help="patterns of the configs[4] leg, WHOLE JOB (strong scaling)")
    ap.add_argument("--mem-steps", type=int, default=3)
    ap.add_argument("--mem-cpu-sample", type=int, default=200_000)
    ap.add_argument("--no-cfg4", action="store_true", help="skip the configs[3] leg (3 Gbp index, 1 B queries sharded over the GPUs)")
    ap.add_argument("--cfg4-mbp", type=float, default=3000.0, help="reference length of the configs[3] leg (Mbp)")
    ap.add_argument("--cfg4-queries", type=int, default=1_000_000_000, help="queries of the configs[3] leg, WHOLE JOB (strong scaling)")
    ap.add_argument("--cfg4-chunk", type=int, default=125_000_000, help="queries per resident chunk of the configs[3] leg")
    ap.add_argument("--cfg4-steps", type=int, default=3, help="timed passes over every chunk of the configs[3] leg")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def workload_name(args):
    return "cfg2: %d x %d-mers sampled from a %g Mbp synthetic linear reference (seed 2), order-128 index (k=16, 3 doubling steps)" % (
        args.queries, args.pattern_length, args.ref_mbp)


def build_or_load_index(args, rank, world, barrier, device=None):
    """The index of configs[1].  With a GPU at hand every rank builds its own copy on its device
    (gcsa_b200_build_linear, about a second for 100 Mbp); otherwise -- or with --host-builder / --index-cache -- rank 0
    builds it on the host (builder.cpp, ~55 s) and shares it through /dev/shm.  Both builders emit identical arrays
    (tests/test_linear_builder.py)."""
    from gcsa2_b200 import synth
    from gcsa2_b200.builder import build_index, build_linear
    from gcsa2_b200.flat import FlatGCSA
    L = int(args.ref_mbp * 1_000_000)
    seq = synth.random_sequence(L, seed=2)
    if device is not None and not args.host_builder and not args.index_cache:
        t0 = time.time()
        flat, _ = build_linear(seq, k=16, doubling_steps=3, node_len=32, device=device)
        return seq, flat, time.time() - t0
    shared = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(),
                          "gcsa2_b200_bench_%d_%d.npz" % (L, os.getppid() if world > 1 else os.getpid()))
    flat = None
    t0 = time.time()
    if rank == 0:
        if args.index_cache and os.path.exists(args.index_cache):
            flat = FlatGCSA.load(args.index_cache)
        else:
            flat, _, _ = build_index(synth.linear_graph(seq, node_len=32), 16, 3)
            if args.index_cache:
                flat.save(args.index_cache)
        if world > 1:
            flat.save(shared)
    barrier()
    if rank != 0:
        flat = FlatGCSA.load(shared)
    barrier()
    if rank == 0 and world > 1 and os.path.exists(shared):
        os.remove(shared)
    return seq, flat, time.time() - t0


def make_patterns(seq, n, length, seed):
    from gcsa2_b200 import synth
    chars = np.empty(n * length, dtype=np.uint8)
    step = 1_000_000
    for i, q0 in enumerate(range(0, n, step)):
        m = min(step, n - q0)
        c, _ = synth.patterns_from_sequence(seq, m, length, seed=seed * 1000 + i)
        chars[q0 * length:(q0 + m) * length] = c
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
    return chars, offsets


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_engine(flat):
    """The CPU implementation to time: the REFERENCE's own sources (oracle/_ref/libgcsa2_ref.so: jltsiren/gcsa2
    src/*.cpp unmodified, compiled against the SDSL shim because sdsl-lite is not available) running its own
    GCSA::find over the same index arrays; if that library is absent, the C restatement (oracle/gcsa_oracle.c)."""
    from oracle import reference as ref
    if ref.available():
        return ref.ReferenceIndex.from_flat(flat), "reference", ref.lib().ref_max_threads()
    from oracle import oracle as orc
    return orc.OracleGCSA(flat), "port", orc.lib().oracle_max_threads()


def cpu_baseline(flat, chars, offsets, length, sample, threads=None):
    """benchmark/query_gcsa.cpp:88-103's loop over GCSA::find, OpenMP over queries like src/algorithms.cpp:113,
    on the host cores."""
    engine, kind, max_threads = cpu_engine(flat)
    threads = threads or max_threads
    n = min(sample, len(offsets) - 1)
    c, o = chars[:n * length], offsets[:n + 1]
    engine.find_batch(c[:length * min(n, 20000)], o[:min(n, 20000) + 1], threads=threads)       # warm the caches
    best = None
    for _ in range(2):
        _, _, secs = engine.find_batch(c, o, threads=threads)
        best = secs if best is None else min(best, secs)
    what = ("reference sources + SDSL shim" if kind == "reference" else "C restatement of the reference")
    return engine, {"value": n / best, "unit": UNIT, "cores": threads, "kind": kind,
                    "sample": "%d of the same %d-mers, best of 2, %d OpenMP threads (schedule dynamic,4096); %s" % (n, length, threads, what),
                    "seconds": best}


def cfg3_fixture(args, rank, world, barrier):
    """The index of BASELINE.json configs[2] and configs[4]: a synthetic variation graph (1 % SNP bubbles over a
    random backbone, seed 3), order 128, with its LCP array.  A graph needs the host builder (rank 0, ~30 s for 50 Mbp);
    the other ranks read the arrays from /dev/shm."""
    from gcsa2_b200 import synth
    from gcsa2_b200.builder import build_index
    from gcsa2_b200.flat import FlatGCSA, FlatLCP
    L = int(args.locate_mbp * 1_000_000)
    seq = synth.random_sequence(L, seed=3)
    graph, sites, alt = synth.snp_graph(seq, seed=3, snp_rate=0.01)
    base = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(),
                        "gcsa2_b200_bench_cfg3_%d_%d" % (L, os.getppid() if world > 1 else os.getpid()))
    t0 = time.time()
    flat = lcp = None
    if rank == 0:
        flat, lcp, _ = build_index(graph, 16, 3)
        if world > 1:
            flat.save(base + ".npz")
            np.savez(base + "_lcp.npz", header=np.array([lcp.size, lcp.branching, lcp.levels], dtype=np.uint64), offsets=lcp.offsets, data=lcp.data)
    barrier()
    if rank != 0:
        flat = FlatGCSA.load(base + ".npz")
        z = np.load(base + "_lcp.npz")
        lcp = FlatLCP(size=int(z["header"][0]), branching=int(z["header"][1]), levels=int(z["header"][2]), offsets=z["offsets"], data=z["data"])
    barrier()
    if rank == 0 and world > 1:
        for name in (base + ".npz", base + "_lcp.npz"):
            if os.path.exists(name):
                os.remove(name)
    return {"seq": seq, "sites": sites, "alt": alt, "flat": flat, "lcp": lcp, "build_s": time.time() - t0, "length": L}


def locate_leg(args, rank, world, local, barrier, dist, torch, fixture):
    """The second half of BASELINE.json's metric: locate() positions/s on configs[2] -- 64-mers sampled from walks
    through a synthetic variation graph (1 % SNP bubbles, order 128), find() then locate(range) as a CSR of sorted
    distinct positions (GCSA::locate, src/gcsa.cpp:827-842).  Device-resident (CUDA events), end to end through
    gcsa_b200_locate_host, and the reference's own locate() on the host cores (rank 0)."""
    from gcsa2_b200 import GCSA, synth
    n, length = args.locate_queries, 64
    seq, sites, alt, flat, build_s = fixture["seq"], fixture["sites"], fixture["alt"], fixture["flat"], fixture["build_s"]
    chars = np.empty(n * length, dtype=np.uint8)
    for i, q0 in enumerate(range(0, n, 1_000_000)):
        m = min(1_000_000, n - q0)
        c, _ = synth.patterns_from_snp_graph(seq, sites, alt, m, length, seed=7000 + 100 * rank + i)
        chars[q0 * length:(q0 + m) * length] = c
    index = GCSA(flat, device=local, kmer_table_k=min(14, args.kmer_table_k))     # (4^14 entries: most 14-mers of this graph are one path node)
    stream = torch.cuda.current_stream()
    d_chars = torch.from_numpy(chars).cuda()
    d_sp = torch.empty(n, dtype=torch.int64, device="cuda"); d_ep = torch.empty_like(d_sp)
    # find() of the 64-mers: timed on its own (not part of the locate metric) -- the k-mer form with the chain kernel
    for _ in range(3):
        index.find_fixed_device(d_chars, length, n, d_sp, d_ep, stream.cuda_stream)
    torch.cuda.synchronize(); barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(3):
        index.find_fixed_device(d_chars, length, n, d_sp, d_ep, stream.cuda_stream)
    f1.record(stream)
    torch.cuda.synchronize()
    find_ms = f0.elapsed_time(f1) / 3
    d_cnt = torch.empty(n, dtype=torch.int64, device="cuda")
    index.count_device(d_sp, d_ep, n, d_cnt, stream.cuda_stream)
    torch.cuda.synchronize()
    total = int(d_cnt.sum().item())
    d_offs = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    d_vals = torch.empty(total + 16, dtype=torch.int64, device="cuda")
    got = [0]

    def step():
        got[0] = index.locate_device(d_sp, d_ep, n, d_offs, d_vals, total + 16, stream.cuda_stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(); barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize(); barrier()
    ms = ev0.elapsed_time(ev1) / args.steps

    # end to end: ranges in pinned host memory -> CSR in pinned host memory
    h_sp = d_sp.cpu().pin_memory(); h_ep = d_ep.cpu().pin_memory()
    h_offs = torch.empty(n + 1, dtype=torch.int64).pin_memory(); h_vals = torch.empty(total + 16, dtype=torch.int64).pin_memory()
    e2e_got = [0]

    def step_e2e():
        e2e_got[0] = index.locate_into_host_raw(h_sp.data_ptr(), h_ep.data_ptr(), n, h_offs.data_ptr(), h_vals.data_ptr(), total + 16)

    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e2e_ms = 1000.0 * (time.perf_counter() - t0) / args.steps
    sp = h_sp.numpy().view(np.uint64); ep = h_ep.numpy().view(np.uint64)
    offs = h_offs.numpy().view(np.uint64); vals = h_vals.numpy().view(np.uint64)[:e2e_got[0]]
    same = bool(e2e_got[0] == got[0] and (d_offs.cpu().numpy().view(np.uint64) == offs).all()
                and (d_vals[:got[0]].cpu().numpy().view(np.uint64) == vals).all())

    positions = got[0]
    if dist is not None:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        c = torch.tensor([positions], dtype=torch.int64, device="cuda")
        dist.all_reduce(c)
        positions = int(c[0])
    out = {"metric": "locate_positions_per_sec", "value": positions / (ms / 1000.0), "unit": "positions/s", "ms_per_step": ms,
           "positions": positions, "ranges": n * world,
           "config": {"workload": "cfg3: %d x 64-mers per GPU from walks through a %g Mbp backbone with 1 %% SNP bubbles (seed 3), order-128 index; "
                                  "find() then locate(range) -> CSR of sorted distinct positions" % (n, args.locate_mbp),
                      "index": {"path_nodes": index.size(), "edges": index.edgeCount(), "device_bytes": index.deviceBytes()}},
           "e2e": {"value": positions / (e2e_ms / 1000.0), "unit": "positions/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": int(16 * n), "d2h_bytes_per_step": int(8 * (n + 1) + 8 * got[0]),
                   "api": "gcsa_b200_locate_into_host (pinned host buffers, chunked H2D/locate/D2H pipeline)", "matches_device_leg": same},
           "setup": {"index_build_s": build_s},
           "find": {"ms_per_step": find_ms, "value": n / (find_ms / 1000.0), "unit": "queries/s per GPU (rank 0)",
                    "note": "find() of the same 64-mers, device-resident, timed separately: find_chain_kernel (table probe, then one jump entry or one "
                            "sector per round) + find_kernel over its work list; kmer_table_k = %d" % min(14, args.kmer_table_k)}}
    if rank == 0:
        # SURVEY.md 8(d), locate: per located node one probe of the locate table (64 B), per range 16 B in and 8 B of
        # offsets out, per position 8 B out.  (locate_small_count / fill read the table once per node: every 64-mer of
        # this workload is a range of at most a few nodes with direct entries.)
        peak, peak_src = peaks()
        nodes = float((d_ep - d_sp + 1).clamp(min=0).sum().item())
        algorithmic = 64.0 * nodes + 24.0 * n + 8.0 * got[0]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_find_cfg2.json")
        if os.path.exists(tpath) and n == 10_000_000 and args.locate_mbp == 50.0:
            with open(tpath) as f:
                entry = json.load(f).get("locate_cfg3")
            traffic = float(entry["dram_bytes_per_launch"]) if entry else None
        out["roofline"] = {"bound": "hbm", "achieved": algorithmic / (ms / 1000.0) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": algorithmic / (ms / 1000.0) / 1e9 / peak, "traffic": traffic,
                           "dram_frac": (traffic / (ms / 1000.0) / 1e9 / peak if traffic else None),
                           "kernel": "locate_small_count_kernel + locate_small_fill_kernel", "peak_source": peak_src,
                           "probe_ceiling": PROBE_CEILING, "probe_rate": nodes / (ms / 1000.0),
                           "probe_frac": (nodes / PROBE_CEILING + (24.0 * n + 8.0 * got[0]) / (peak * 1e9)) / (ms / 1000.0),
                           "accounting": "64 B per located path node (one locate-table entry) + 16 B per range in + 8 B of offsets and 8 B per position out; "
                                         "probe_frac as for find(): time at the HBM random-access rate for the probes + time at the copy peak for the streams, over the time taken"}
    if rank == 0 and not args.no_cpu_baseline:
        engine, kind, threads = cpu_engine(flat)
        m = min(n, 50_000 * threads)
        roffs, rvals, secs = engine.locate_batch(sp[:m], ep[:m], threads=threads)
        k = int(roffs[m])
        out["cpu_baseline"] = {"value": k / secs, "unit": "positions/s", "cores": threads, "kind": kind,
                               "sample": "locate() of the first %d ranges, %d OpenMP threads (schedule dynamic,256)" % (m, threads), "seconds": secs,
                               "parity_on_sample": bool((offs[:m + 1] == roffs).all() and (vals[:k] == rvals).all())}
    if rank == 0 and not args.no_wide_locate:
        out["wide_ranges"] = wide_locate(args, index, fixture, torch)
    index.close()
    return out


def wide_locate(args, index, fixture, torch):
    """locate() of the ranges of SHORT patterns on the configs[2] index (rank 0, device-resident): tens to thousands of path
    nodes per range, sorted and deduplicated in registers by a warp or a block (locate_medium_kernel) instead of one
    thread per range.  One line per pattern length, each checked against the CPU engine on a sample."""
    from gcsa2_b200 import synth
    seq, sites, alt, flat = fixture["seq"], fixture["sites"], fixture["alt"], fixture["flat"]
    stream = torch.cuda.current_stream()
    scale = max(args.locate_queries / 10_000_000.0, 0.0005)
    lines = []
    for plen, base_n in ((10, 1_000_000), (8, 200_000), (6, 20_000)):
        nq = max(int(base_n * scale), 64)
        pchars, poffsets = synth.patterns_from_snp_graph(seq, sites, alt, nq, plen, seed=600 + plen)
        psp, pep = index.find_batch(pchars, poffsets)
        d_sp = torch.from_numpy(psp.view(np.int64)).cuda(); d_ep = torch.from_numpy(pep.view(np.int64)).cuda()
        d_cnt = torch.empty(nq, dtype=torch.int64, device="cuda")
        index.count_device(d_sp, d_ep, nq, d_cnt, stream.cuda_stream); torch.cuda.synchronize()
        total = int(d_cnt.sum().item())
        d_offs = torch.empty(nq + 1, dtype=torch.int64, device="cuda"); d_vals = torch.empty(total + 16, dtype=torch.int64, device="cuda")
        got = [0]

        def step():
            got[0] = index.locate_device(d_sp, d_ep, nq, d_offs, d_vals, total + 16, stream.cuda_stream)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            step()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        line = {"pattern_length": plen, "ranges": nq, "path_nodes_per_range": float((pep - psp + 1).astype(np.float64).mean()),
                "positions": got[0], "ms_per_step": ms, "value": got[0] / (ms / 1000.0), "unit": "positions/s"}
        if not args.no_cpu_baseline:
            engine, kind, threads = cpu_engine(flat)
            m = min(nq, max(64, 4_000_000 // max(1, int(line["path_nodes_per_range"]))))
            roffs, rvals, secs = engine.locate_batch(psp[:m], pep[:m], threads=threads)
            k = int(roffs[m])
            line["cpu_baseline"] = {"value": k / secs, "unit": "positions/s", "cores": threads, "kind": kind, "sample": "the first %d ranges" % m,
                                    "parity_on_sample": bool(got[0] == total and (d_offs[:m + 1].cpu().numpy().view(np.uint64) == roffs).all()
                                                             and (d_vals[:k].cpu().numpy().view(np.uint64) == rvals).all())}
        lines.append(line)
        del d_sp, d_ep, d_cnt, d_offs, d_vals
    return lines


def mem_leg(args, rank, world, local, barrier, dist, torch, fixture):
    """BASELINE.json configs[4]: mixed-length 16-256 bp patterns through the MEM-style scan (GCSA::LF + LCPArray::parent,
    include/gcsa/gcsa.h:155-162, src/lcp.cpp:276-301) on the configs[2] graph and its LCP array -- the warp-divergence
    stress.  STRONG scaling: the job is --mem-patterns patterns for every N; they are generated on the device (the
    same batch on every rank), ordered by length and dealt round-robin (gcsa2_b200.dist.shard_patterns_by_length's rule),
    so every rank scans the same mix of lengths.  value = patterns of the whole job / max over ranks of the scan time."""
    from gcsa2_b200 import GCSA, LCPArray, mem_device, synth
    flat, flcp = fixture["flat"], fixture["lcp"]
    total = args.mem_patterns
    d_seq = torch.from_numpy(fixture["seq"]).cuda()
    all_chars, all_offsets = synth.device_mixed_length_patterns(d_seq, fixture["sites"], fixture["alt"], total, 16, 256, seed=5, error_rate=0.01)
    d_chars, d_offsets, ids = synth.device_shard_by_length(all_chars, all_offsets, rank, world)
    total_bytes = int(all_chars.numel())
    del all_chars, all_offsets, d_seq
    n = int(ids.numel())
    index = GCSA(flat, device=local, kmer_table_k=0, walk_table=0, jump_table=False)      # the scan uses the fused blocks only
    lcp = LCPArray(flcp, device=local)
    stream = torch.cuda.current_stream()
    d_out = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    cap = 16 * n
    d_matches = torch.empty((cap, 4), dtype=torch.int64, device="cuda")
    got = [0]

    def step():
        got[0] = mem_device(index, lcp, d_chars, d_offsets, n, d_out, d_matches, cap, stream.cuda_stream)

    for _ in range(2):
        step()
    torch.cuda.synchronize(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.mem_steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(); barrier()
    ms = e0.elapsed_time(e1) / args.mem_steps
    matches, patterns, my_bytes = got[0], n, int(d_offsets[-1].item())
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        cnt = torch.tensor([patterns, matches], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        patterns, matches = int(cnt[0]), int(cnt[1])
    out = {"metric": "mem_scan_patterns_per_sec", "value": patterns / (ms / 1000.0), "unit": "patterns/s", "scaling": "strong", "n_gpus": world,
           "ms_per_step": ms, "patterns": patterns, "matches": matches, "pattern_bytes": total_bytes,
           "config": {"workload": "cfg5: %d patterns, lengths uniform in 16..256, sampled from walks through the %g Mbp 1 %% SNP graph (seed 3) with 1 %% "
                                  "substitutions (device generator, seed 5); MEM-style scan (LF while the range is not empty, else report and parent()); "
                                  "dealt round-robin in length order to %d GPU(s), index + LCP array replicated" % (total, args.locate_mbp, world),
                      "index": {"path_nodes": index.size(), "device_bytes": index.deviceBytes(), "lcp_levels": int(flcp.levels)},
                      "steps": args.mem_steps}}
    if rank == 0:
        # Work of the reference loop on this rank's share: one LF step per pattern character (2 rank probes on B_c and 2 on
        # `edges` in the reference; here two fused sectors at most) and one parent() per reported match; 32 B per match out.
        peak, peak_src = peaks()
        algorithmic = 128.0 * my_bytes + 128.0 * got[0] + 32.0 * got[0] + my_bytes + 8.0 * n
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_find_cfg2.json")
        if args.locate_mbp == 50.0 and os.path.exists(tpath):
            with open(tpath) as f:
                entry = json.load(f).get("mem_cfg5")
            traffic = float(entry["dram_bytes_per_launch"]) / float(entry["patterns_per_launch"]) * n if entry else None
        out["roofline"] = {"bound": "hbm", "achieved": algorithmic / (ms / 1000.0) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": algorithmic / (ms / 1000.0) / 1e9 / peak, "traffic": traffic,
                           "dram_frac": (traffic / (ms / 1000.0) / 1e9 / peak if traffic else None), "kernel": "mem_kernel<2,false>", "peak_source": peak_src,
                           "accounting": "per GPU (rank 0): 2 fused sectors (2 x 64 B) per pattern character, 2 x 64 B per parent(), 32 B per match written, "
                                         "the pattern bytes and 8 B of offsets per pattern; traffic = recorded ncu DRAM bytes per pattern x this rank's patterns.  "
                                         "The index (84 MB of fused blocks, 58 MB of LCP tree) is larger than what the L2 keeps (hit rate 35 %), but the scan is "
                                         "bound by divergence (15.6 of 32 lanes active per instruction) and load latency, not by HBM throughput: "
                                         "profiles/r02_ncu_mem_kernel.txt, profiles/r02_mem_scan_variants.txt"}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        threads = orc.lib().oracle_max_threads()
        m = min(n, args.mem_cpu_sample)
        h_off = d_offsets[:m + 1].cpu().numpy().astype(np.uint64)
        h_chars = d_chars[:int(h_off[m])].cpu().numpy()
        eoffs, evals, secs = orc.mem_batch(orc.OracleGCSA(flat), orc.OracleLCP(flcp), h_chars, h_off, threads=threads)
        k = int(eoffs[m])
        moffs = d_out[:m + 1].cpu().numpy().view(np.uint64); mvals = d_matches[:k].cpu().numpy().view(np.uint64)
        out["cpu_baseline"] = {"value": m / secs, "unit": "patterns/s", "cores": threads, "kind": "port", "seconds": secs,
                               "sample": "the first %d patterns of rank 0's share, %d OpenMP threads; the scan loop is this repository's (the reference ships "
                                         "LF and parent but no driver), run over the C restatement of both" % (m, threads),
                               "parity_on_sample": bool((moffs == eoffs).all() and (mvals.reshape(-1, 4) == evals.reshape(-1, 4)).all())}
    index.close(); lcp.close()
    return out


def cfg4_traffic(args, queries):
    """Recorded DRAM bytes of the find() kernels on the 3 Gbp index (ncu --set full of a 10 M query launch, per query),
    scaled to `queries`; None for any other configuration."""
    path = os.path.join(ROOT, "profiles", "traffic_find_cfg2.json")
    if args.cfg4_mbp != 3000.0 or args.kmer_table_k != 16 or not os.path.exists(path):
        return None
    with open(path) as f:
        entry = json.load(f).get("cfg4_3gbp")
    return float(entry["dram_bytes_per_launch"]) / float(entry["queries_per_launch"]) * queries if entry else None


def cfg4_leg(args, rank, world, local, barrier, dist, torch):
    """BASELINE.json configs[3]: 1 B 32-mers on a 3 Gbp linear-path order-128 index, the batch sharded across the
    GPUs, the index replicated -- STRONG scaling: the job is the same 1 B queries for every N.  Nothing of it touches
    the host: the reference is generated on the device (counter-based generator, identical on every rank), the index
    is built there (gcsa_b200_build_linear), and so are the patterns, one resident chunk of --cfg4-chunk queries at a
    time (a rank owns the chunks c with c % world == rank; 1 B x 32 B would not fit next to the index on one GPU).
    Timed: --cfg4-steps passes of find() over every chunk (CUDA events on the launching stream), chunk generation
    excluded; value = queries of the whole job / max over ranks of the summed time of one pass."""
    from gcsa2_b200 import GCSA, synth
    from gcsa2_b200.builder import build_linear
    L, total, chunk, length = int(args.cfg4_mbp * 1_000_000), args.cfg4_queries, args.cfg4_chunk, 32
    n_chunks = (total + chunk - 1) // chunk
    mine = [c for c in range(n_chunks) if c % world == rank]
    t0 = time.time()
    seq = synth.device_sequence(L, seed=4)
    built = build_linear(seq, k=16, doubling_steps=3, node_len=32, device=local, raw=True)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    t0 = time.time()
    index = GCSA(built, device=local, kmer_table_k=args.kmer_table_k, walk_table=0, two_step=False, fused_table=False)     # find() only: no locate tables
    torch.cuda.synchronize()
    create_s = time.time() - t0
    flat = built.flat() if (rank == 0 and not args.no_cpu_baseline) else None
    built.free()
    stream = torch.cuda.current_stream()
    d_sp = torch.empty(min(chunk, total), dtype=torch.int64, device="cuda"); d_ep = torch.empty_like(d_sp)
    ms_pass, found, queries, sample = 0.0, 0, 0, None
    for c in mine:
        m = min(chunk, total - c * chunk)
        d_chars = synth.device_patterns(seq, m, length, seed=4000 + c)
        for _ in range(2):
            index.find_fixed_device(d_chars, length, m, d_sp, d_ep, stream.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.cfg4_steps):
            index.find_fixed_device(d_chars, length, m, d_sp, d_ep, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms_pass += e0.elapsed_time(e1) / args.cfg4_steps
        found += int(((d_sp[:m] + 1) <= (d_ep[:m] + 1)).sum().item()); queries += m
        if sample is None and rank == 0:
            k = min(m, 1_000_000)
            sample = (d_chars[:k * length].cpu().numpy(), d_sp[:k].cpu().numpy().view(np.uint64).copy(), d_ep[:k].cpu().numpy().view(np.uint64).copy())
        del d_chars
    barrier()
    # work counters of the kernel on the parity sample (untimed, stats variant)
    st = None
    if sample is not None:
        k = sample[1].size
        _, _, st = index.find_fixed_batch(sample[0], length, stats=True)
    info = {"path_nodes": index.size(), "edges": index.edgeCount(), "order": index.order(), "device_bytes": index.deviceBytes(),
            "kmer_table_k": index.kmerTableK(), "fused_table": index.fusedTable(), "two_step": index.twoStep(), "jump_k": index.jumpK()}
    index.close()
    del seq, d_sp, d_ep
    torch.cuda.empty_cache()
    if dist is not None:
        t = torch.tensor([ms_pass], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_pass = float(t[0])
        cnt = torch.tensor([queries, found], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)                                          # NCCL: the gather of the result counters
        queries, found = int(cnt[0]), int(cnt[1])
    out = {"metric": METRIC, "value": queries / (ms_pass / 1000.0), "unit": UNIT, "scaling": "strong", "n_gpus": world,
           "ms_per_pass": ms_pass, "queries": queries, "found": found,
           "config": {"workload": "cfg4: %d x 32-mers sampled from a %g Mbp synthetic linear reference (device generator, seed 4), order-128 index "
                                  "(k=16, 3 doubling steps) built on the device; %d resident chunks of %d queries dealt round-robin to %d GPU(s), index replicated" % (
                                      total, args.cfg4_mbp, n_chunks, chunk, world),
                      "index": info, "steps_per_chunk": args.cfg4_steps},
           "setup": {"sequence_and_build_s": build_s, "index_create_s": create_s}}
    if rank == 0 and st is not None:
        peak, peak_src = peaks()
        k = sample[1].size
        per_query = 64.0 * (st["sector_probes"] + st["table_hits"]) / k + length + 16
        secs = ms_pass / 1000.0
        rank0_queries = sum(min(chunk, total - c * chunk) for c in mine)
        out["roofline"] = {"bound": "hbm", "achieved": per_query * rank0_queries / secs / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": per_query * rank0_queries / secs / 1e9 / peak, "traffic": cfg4_traffic(args, rank0_queries),
                           "dram_frac": (cfg4_traffic(args, rank0_queries) / secs / 1e9 / peak if cfg4_traffic(args, rank0_queries) else None), "kernel": "find_fast_kernel<false,false,4> (+ find_quad_kernel, find_kernel<false,4,false,true> for the work lists)",
                           "peak_source": peak_src, "probes_per_query": (st["sector_probes"] + st["table_hits"]) / k,
                           "probe_ceiling": PROBE_CEILING, "probe_ceiling_source": PROBE_CEILING_SOURCE,
                           "probe_rate": (st["sector_probes"] + st["table_hits"]) / k * rank0_queries / secs,
                           "probe_frac": ((st["sector_probes"] + st["table_hits"]) / k * rank0_queries / PROBE_CEILING + (length + 16.0) * rank0_queries / (peak * 1e9)) / secs,
                           "lf_steps_per_query": st["lf_steps"] / k,
                           "accounting": "per GPU (rank 0): 64 B per distinct probe executed + |P| + 16 B I/O per query, over the time of one pass"}
    if rank == 0 and flat is not None:
        t0 = time.time()
        engine, kind, threads = cpu_engine(flat)
        load_s = time.time() - t0
        k = sample[1].size
        offsets = np.arange(k + 1, dtype=np.uint64) * np.uint64(length)
        csp, cep, secs = engine.find_batch(sample[0], offsets, threads=threads)
        out["cpu_baseline"] = {"value": k / secs, "unit": UNIT, "cores": threads, "kind": kind, "seconds": secs,
                               "sample": "the first %d queries of chunk 0, %d OpenMP threads; %s" % (
                                   k, threads, "reference sources + SDSL shim" if kind == "reference" else "C restatement of the reference"),
                               "parity_on_sample": bool((csp == sample[1]).all() and (cep == sample[2]).all()), "index_load_s": load_s}
    return out


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(args, index):
    """dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full capture of this
    very configuration (profiles/traffic_find_cfg2.json, one entry per k-mer table size); None for any other."""
    path = os.path.join(ROOT, "profiles", "traffic_find_cfg2.json")
    default = (args.ref_mbp == 100.0 and args.queries == 10_000_000 and args.pattern_length == 32 and not index.twoStep())
    if default and os.path.exists(path):
        with open(path) as f:
            entry = json.load(f).get("k%d%s" % (index.kmerTableK(), "f" if index.fusedTable() else ""))
        if entry:
            return float(entry["dram_bytes_per_launch"])
    return None


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (restated in oracle/; the reference itself cannot be
    built here because sdsl-lite is absent), all host threads, a bounded sample per step."""
    if rank != 0:
        return
    device = None
    try:                                    # the fixture may be built on a GPU if there is one; the timed loop below is CPU only
        import torch
        if torch.cuda.is_available():
            device = 0
    except Exception:
        device = None
    seq, flat, build_s = build_or_load_index(args, 0, 1, lambda: None, device=device)
    engine, kind, threads = cpu_engine(flat)
    sample = args.cpu_sample or min(args.queries, 200_000 * threads)
    # the first `sample` queries of the very batch rank 0 of the GPU arm searches (seed 100)
    chars, offsets = make_patterns(seq, args.queries, args.pattern_length, seed=100)
    chars, offsets = chars[:sample * args.pattern_length], offsets[:sample + 1]
    times = []
    for i in range(args.warmup + args.steps):
        _, _, secs = engine.find_batch(chars, offsets, threads=threads)
        if i >= args.warmup:
            times.append(secs)
    ms = 1000.0 * float(np.mean(times))
    value = sample / (ms / 1000.0)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args), "step": "bounded sample per step: the first %d queries of the GPU arm's rank-0 batch (seed 100)" % sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": "%d queries per step, %d OpenMP threads; %s" % (
                                 sample, threads, "reference sources + SDSL shim" if kind == "reference" else "C restatement")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "index_build_s": build_s}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier(device_ids=[local])

    from gcsa2_b200 import GCSA
    seq, flat, build_s = build_or_load_index(args, rank, world, barrier, device=local)
    n, length = args.queries, args.pattern_length
    chars, offsets = make_patterns(seq, n, length, seed=100 + rank)

    t0 = time.time()
    options = dict(device=local, kmer_table_k=args.kmer_table_k, two_step=(None if args.two_step < 0 else bool(args.two_step)))
    create_note = None
    try:
        index = GCSA(flat, fused_table=(None if args.fused_table < 0 else bool(args.fused_table)), **options)
    except Exception as exc:                                        # e.g. not enough free HBM for the 16-byte entries
        create_note = "fused table failed (%s), 8-byte entries instead" % exc
        index = GCSA(flat, fused_table=False, **options)
    create_s = time.time() - t0

    # ---- device-resident leg ----
    d_chars = torch.from_numpy(chars).cuda()
    d_sp = torch.empty(n, dtype=torch.int64, device="cuda"); d_ep = torch.empty_like(d_sp)
    stream = torch.cuda.current_stream()

    def step_device():
        index.find_fixed_device(d_chars, length, n, d_sp, d_ep, stream.cuda_stream)

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize(); barrier()
    import ctypes
    from gcsa2_b200 import capi
    fast_launches = capi.lib().gcsa_b200_internal_fast_launches
    fast_launches.restype = ctypes.c_ulonglong
    fast_before = fast_launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        # The timed region is a few milliseconds and nvidia-smi reports every 100 ms: the same step keeps the GPU under the
        # same load (untimed) until two samples are in, and again after the timed steps until one more is, so that the
        # clocks line describes the load the timed steps ran under.
        t_load = time.perf_counter()
        while len(sampler.lines) < 2 and time.perf_counter() - t_load < 2.0:
            for _ in range(20):
                step_device()
            torch.cuda.synchronize()
    fast_before = fast_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    torch.cuda.synchronize(); barrier()
    ms_total = ev0.elapsed_time(ev1)
    fast_after = fast_launches()
    if rank == 0:
        seen, t_load = len(sampler.lines), time.perf_counter()
        while len(sampler.lines) <= seen and time.perf_counter() - t_load < 0.5:
            for _ in range(20):
                step_device()
            torch.cuda.synchronize()
    # kernels of this library launched inside the timed region: a batch in the k-mer form is three launches
    # (find_fast_kernel, find_quad_kernel, find_kernel over the work list), any other batch one (find_kernel)
    fast_steps = int(fast_after - fast_before)
    gpu_launches = 3 * fast_steps + (args.steps - fast_steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps

    # results of the last step: counters
    sp = d_sp.cpu().numpy().view(np.uint64); ep = d_ep.cpu().numpy().view(np.uint64)
    found = int(np.count_nonzero((sp + np.uint64(1)) <= (ep + np.uint64(1))))

    # ---- secondary workload of configs[1]: uniform random 32-mers, which miss after ~log4(N) steps ----
    secondary = None
    try:
        from gcsa2_b200 import synth
        rchars, _ = synth.random_patterns(n, length, seed=900 + rank)
        d_rchars = torch.from_numpy(rchars).cuda()
        d_rsp = torch.empty(n, dtype=torch.int64, device="cuda"); d_rep = torch.empty_like(d_rsp)
        for _ in range(3):
            index.find_fixed_device(d_rchars, length, n, d_rsp, d_rep, stream.cuda_stream)
        torch.cuda.synchronize(); barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        for _ in range(args.steps):
            index.find_fixed_device(d_rchars, length, n, d_rsp, d_rep, stream.cuda_stream)
        r1.record(stream)
        torch.cuda.synchronize(); barrier()
        rms = r0.elapsed_time(r1) / args.steps
        rsp = d_rsp.cpu().numpy().view(np.uint64); rep = d_rep.cpu().numpy().view(np.uint64)
        secondary = {"workload": "%d uniform random %d-mers per GPU (early exit)" % (n, length), "ms_per_step": rms,
                     "value": n / (rms / 1000.0), "unit": UNIT,
                     "found": int(np.count_nonzero((rsp + np.uint64(1)) <= (rep + np.uint64(1))))}
        del d_rchars, d_rsp, d_rep
    except Exception as exc:                                        # the primary line must survive a failure here
        secondary = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # ---- end-to-end leg: pinned host buffers through gcsa_b200_find_host ----
    h_chars = torch.from_numpy(chars).pin_memory()
    h_sp = torch.empty(n, dtype=torch.int64).pin_memory(); h_ep = torch.empty(n, dtype=torch.int64).pin_memory()

    # Host-side 2-bit packing before the H2D copy (gcsa2_b200/csrc/pack.cpp): the host entry point shares the batch
    # between a raw-copy thread and the packing threads (GCSA_B200_HOST_PACK=0 forbids packing, =N sets the threads).
    if world > 1 and not os.environ.get("GCSA_B200_HOST_PACK_THREADS"):
        os.environ["GCSA_B200_HOST_PACK_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))

    def step_e2e():
        index.find_fixed_host_raw(h_chars.data_ptr(), length, n, h_sp.data_ptr(), h_ep.data_ptr())

    e2e_note = None
    try:
        for _ in range(max(args.warmup, 3)):
            step_e2e()
    except Exception as exc:                                        # the shared raw / packed pipeline failed: raw copies only
        e2e_note = "host packing disabled after: %s" % exc
        os.environ["GCSA_B200_HOST_PACK"] = "0"
        for _ in range(max(args.warmup, 3)):
            step_e2e()
    torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = 1000.0 * (time.perf_counter() - t0) / args.steps
    barrier()
    e2e_same = bool((h_sp.numpy().view(np.uint64) == sp).all() and (h_ep.numpy().view(np.uint64) == ep).all())
    pack_env = os.environ.get("GCSA_B200_HOST_PACK") or "auto"
    c_packed, c_total = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    capi.lib().gcsa_b200_internal_pack_share(ctypes.byref(c_packed), ctypes.byref(c_total))     # of the last e2e step
    pack_share = c_packed.value / max(1, c_total.value)
    pack_threads = 0
    if pack_env != "0":
        pack_threads = int(pack_env) if pack_env != "auto" else int(os.environ.get("GCSA_B200_HOST_PACK_THREADS") or os.environ.get("OMP_NUM_THREADS") or os.cpu_count() or 1)

    # ---- work counters for the roofline (untimed; a 1 M sample through the stats kernel) ----
    m = min(n, 1_000_000)
    _, _, st = index.find_fixed_batch(chars[:m * length], length, stats=True)       # the kernels of the k-mer form
    scale = n / m
    # 64 B per distinct probe (the HBM access granule: a missed 32-byte sector costs one 64-byte
    # fetch) + |P| + 16 B of I/O per query.  A probe here is a fused sector, a jump-table entry or a k-mer table
    # entry: each is one random access into a structure far larger than the L2.  (Until the fused table the k-mer
    # table entry was charged its 8 payload bytes only; `achieved_payload` keeps that stricter figure.)
    entry_bytes = 16.0 if index.fusedTable() else 8.0
    payload_bytes = scale * (64.0 * st["sector_probes"] + entry_bytes * st["table_hits"]) + float(n) * (length + 16)
    engine_bytes = scale * 64.0 * (st["sector_probes"] + st["table_hits"]) + float(n) * (length + 16)

    # ---- max over ranks, totals ----
    if dist is not None:
        t = torch.tensor([ms_step, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms = float(t[0]), float(t[1])
        from gcsa2_b200 import dist as gd
        c = gd.all_reduce_counters(gd.find_counters(sp, ep), device="cuda")   # NCCL: gather of the result counters
        total_q, total_found = int(c[0]), int(c[1])
    else:
        total_q, total_found = n, found

    locate = mem = fixture = None
    if not (args.no_locate and args.no_mem):
        try:
            del d_chars, h_chars
            fixture = cfg3_fixture(args, rank, world, barrier)
        except Exception as exc:
            locate = mem = {"error": "cfg3 fixture: %s: %s" % (type(exc).__name__, exc)}
    if fixture is not None and not args.no_locate:
        try:
            locate = locate_leg(args, rank, world, local, barrier, dist, torch, fixture)
        except Exception as exc:                                    # the find() line must survive a failure here
            locate = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if fixture is not None and not args.no_mem:
        try:
            torch.cuda.empty_cache()
            mem = mem_leg(args, rank, world, local, barrier, dist, torch, fixture)
        except Exception as exc:
            mem = {"error": "%s: %s" % (type(exc).__name__, exc)}
    fixture = None

    cfg4 = None
    if not args.no_cfg4:
        try:
            if torch.cuda.get_device_properties(local).total_memory < 150e9:
                raise RuntimeError("needs a GPU with more than 150 GB")
            index.close()                                           # the 3 Gbp index needs the HBM (the accessors below are cached)
            d_sp = d_ep = h_sp = h_ep = None
            torch.cuda.empty_cache()
            cfg4 = cfg4_leg(args, rank, world, local, barrier, dist, torch)
        except Exception as exc:                                    # the find() line must survive a failure here
            cfg4 = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank == 0:
        peak, peak_src = peaks()
        traffic = recorded_traffic(args, index)
        achieved = engine_bytes / (ms_total / args.steps / 1000.0) / 1e9
        line = {
            "metric": METRIC, "value": total_q / (ms_step / 1000.0), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args), "queries_per_gpu": n, "pattern_length": length,
                       "index": {"path_nodes": index.size(), "edges": index.edgeCount(), "order": index.order(),
                                 "device_bytes": index.deviceBytes(), "kmer_table_k": index.kmerTableK(), "fused_table": index.fusedTable(), "two_step": index.twoStep()},
                       "parallelism": "queries sharded across %d GPU(s), index replicated" % world,
                       "l2": "no explicit flush: every step streams %.0f MB of patterns/offsets/results, %s the 126 MB L2" % (
                           n * (length + 16) / 1e6, "more than" if n * (length + 16) > 126e6 else "LESS than (reduced run: not a valid timing)")},
            "found": total_found, "queries": total_q,
            "device_bytes": index.deviceBytes(), "kmer_table_k": index.kmerTableK(),
            "e2e": {"value": total_q / (e2e_ms / 1000.0), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(n * (pack_share * 8 * ((length + 31) // 32) + (1.0 - pack_share) * length)), "d2h_bytes_per_step": int(n * 16),
                    "api": "gcsa_b200_find_fixed_host (pinned host buffers, chunked H2D/kernel/D2H pipeline%s)" % (
                        "; a raw-copy thread and %d packing threads share the batch: %.0f %% of the chunks crossed PCIe 2-bit packed, %d B/query instead of %d" % (
                            pack_threads, 100.0 * pack_share, 8 * ((length + 31) // 32), length) if pack_threads > 0 else ""),
                    "host_pack": {"policy": pack_env, "threads": pack_threads, "packed_chunks": c_packed.value, "chunks": c_total.value},
                    "matches_device_leg": e2e_same},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "dram_frac": (traffic / (ms_total / args.steps / 1000.0) / 1e9 / peak if traffic else None),
                         "kernel": "find_fast_kernel<false,false,4> (+ find_quad_kernel, find_kernel<false,4,false,true> for the work lists)", "peak_source": peak_src,
                         "bytes_per_launch": engine_bytes,
                         "achieved_payload": payload_bytes / (ms_total / args.steps / 1000.0) / 1e9,
                         "probes_per_query": (st["sector_probes"] + st["table_hits"]) / m,
                         "accounting": "64 B per distinct probe executed (fused sector, jump-table entry or k-mer table entry) + |P| + 16 B I/O per query (SURVEY.md 8(d) units); "
                                       "achieved_payload charges a k-mer table entry its payload only (8 B, 16 B fused), as the lines of profiles/r01_bench_cfg2_*.json did; "
                                       "dram_frac = recorded ncu DRAM bytes of this launch / this run's time / peak: a random probe into tens of GB costs ~128 B of HBM traffic, "
                                       "twice the 64 B the accounting grants it (profiles/r01_random_probe_microbench.txt)",
                         # the roofline that binds a random probe into tens of GB is the HBM random-access rate, not the copy
                         # bandwidth: time at that rate for the probes + time at the copy peak for the streams, over the time taken
                         "probe_ceiling": PROBE_CEILING, "probe_ceiling_source": PROBE_CEILING_SOURCE,
                         "probe_rate": scale * (st["sector_probes"] + st["table_hits"]) / (ms_total / args.steps / 1000.0),
                         "probe_frac": (scale * (st["sector_probes"] + st["table_hits"]) / PROBE_CEILING + float(n) * (length + 16) / (peak * 1e9)) / (ms_total / args.steps / 1000.0),
                         "jump_table_k": index.jumpK(),
                         "lf_steps_per_query": st["lf_steps"] / m, "sector_probes_per_query": st["sector_probes"] / m},
            "clocks": clocks,
            "setup": {"index_build_s": build_s, "index_create_s": create_s},
        }
        if create_note:
            line["setup"]["note"] = create_note
        if e2e_note:
            line["e2e"]["note"] = e2e_note
        if locate is not None:
            line["locate"] = locate
        if mem is not None:
            line["cfg5"] = mem
        if cfg4 is not None:
            line["cfg4"] = cfg4
        if secondary is not None:
            if dist is not None and "value" in secondary:
                secondary["note"] = "rank 0 only"
            line["secondary"] = secondary
        if not args.no_cpu_baseline:
            from oracle import oracle as orc
            threads = orc.lib().oracle_max_threads()
            sample = args.cpu_sample or min(n, 200_000 * threads)
            engine, cb = cpu_baseline(flat, chars, offsets, length, sample, threads)
            csp, cep, _ = engine.find_batch(chars[:m * length], offsets[:m + 1], threads=threads)
            cb["parity_on_sample"] = bool((csp == sp[:m]).all() and (cep == ep[:m]).all())
            if secondary is not None and "value" in secondary:
                k2 = min(n, 200_000)
                csp, cep, _ = engine.find_batch(rchars[:k2 * length], offsets[:k2 + 1], threads=threads)
                secondary["parity_on_sample"] = bool((csp == rsp[:k2]).all() and (cep == rep[:k2]).all())
            line["cpu_baseline"] = cb
            # probes the reference algorithm issues (counted by the C restatement while answering)
            _, _, _, steps_ref, probes_ref = orc.OracleGCSA(flat).find_batch(chars[:m * length], offsets[:m + 1], threads=threads, stats=True)
            ref_bytes = scale * 64.0 * probes_ref + float(n) * (length + 16)
            line["roofline"]["reference_accounting"] = {
                "bytes_per_launch": ref_bytes, "achieved": ref_bytes / (ms_total / args.steps / 1000.0) / 1e9,
                "note": "64 B per distinct rank probe the REFERENCE algorithm issues (4 per LF step, no k-mer table) + |P| + 16 B; the fused layout and the table avoid most of them"}
        print(json.dumps(line), flush=True)

    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
