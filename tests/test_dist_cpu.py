"""The N > 1 path on CPU: world_size-2 gloo, query sharding and the counter all-reduce.
The search itself is stood in for by the CPU oracle (test infrastructure) -- what is under test is
the host-side sharding / gathering logic of gcsa2_b200.dist."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gcsa2_b200 import dist as gd, synth
    from gcsa2_b200.builder import build_index
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = synth.random_sequence(20000, seed=1)
    flat, _, _ = build_index(synth.linear_graph(seq), 16, 1)          # every rank: its own replica
    ora = orc.OracleGCSA(flat)
    chars, offsets = synth.patterns_from_sequence(seq, 1001, 20, seed=3)
    rchars, roffsets = synth.random_patterns(500, 20, seed=4)
    chars = np.concatenate([chars, rchars]); offsets = np.concatenate([offsets, roffsets[1:] + offsets[-1]])
    my_chars, my_offsets, (q0, q1) = gd.shard_patterns(chars, offsets, rank, world)
    sp, ep, _ = ora.find_batch(my_chars, my_offsets)
    total = gd.all_reduce_counters(gd.find_counters(sp, ep))
    all_sp, all_ep = gd.gather_ranges(sp, ep)
    if rank == 0:
        full_sp, full_ep, _ = ora.find_batch(chars, offsets)
        expect = gd.find_counters(full_sp, full_ep)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([
            int((total == expect).all()), int((all_sp == full_sp).all() and (all_ep == full_ep).all()),
            int(total[0]), int(total[1])]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    from gcsa2_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 1001):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_shard_and_gather(tmp_path):
    port = free_port()
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok[0] == 1 and ok[1] == 1 and ok[2] == 1501 and 1001 <= ok[3] <= 1501


def test_shard_by_length_gives_every_rank_the_same_mix():
    """configs[4] (patterns of 16-256 bp): the shards cover the batch exactly once, hold the right characters, and
    carry the same number of characters to within one pattern per rank."""
    from gcsa2_b200.dist import shard_patterns_by_length
    rng = np.random.default_rng(5)
    lengths = rng.integers(16, 257, size=5003)
    lengths[:7] = 0                                                    # empty patterns too
    offsets = np.zeros(lengths.size + 1, dtype=np.uint64); offsets[1:] = np.cumsum(lengths)
    chars = rng.integers(65, 91, size=int(offsets[-1]), dtype=np.uint8)
    for world in (1, 2, 4, 7):
        seen, totals = [], []
        for rank in range(world):
            c, o, ids = shard_patterns_by_length(chars, offsets, rank, world)
            assert o[0] == 0 and len(o) == len(ids) + 1 and (np.diff(ids) > 0).all()
            for j in (0, len(ids) // 2, len(ids) - 1):
                i = int(ids[j])
                assert bytes(c[int(o[j]):int(o[j + 1])]) == bytes(chars[int(offsets[i]):int(offsets[i + 1])])
            seen.append(ids); totals.append(int(o[-1]))
        assert sorted(np.concatenate(seen).tolist()) == list(range(lengths.size))
        assert max(totals) - min(totals) <= 256 * 2


def strong_scaling_worker(rank, world, port, out_dir):
    """The strong-scaling legs of bench.py (configs[3]: chunks of one job dealt round-robin; configs[4]: patterns dealt in
    length order) with the CPU oracle standing in for the kernels: the all-reduced counters must be those of the whole job."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gcsa2_b200 import synth
    from gcsa2_b200.builder import build_index
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = synth.counter_sequence(30_000, seed=4)                        # every rank regenerates the same reference
    graph, sites, alt = synth.snp_graph(seq, seed=3, snp_rate=0.01)
    flat, flcp, _ = build_index(graph, 16, 2)
    ora, olcp = orc.OracleGCSA(flat), orc.OracleLCP(flcp)
    # configs[3]: 7 chunks of 1000 queries (the last one shorter), chunk c belongs to rank c % world
    total, chunk, length = 6500, 1000, 32
    t_seq = torch.from_numpy(seq)
    mine = [c for c in range((total + chunk - 1) // chunk) if c % world == rank]
    counts = torch.zeros(2, dtype=torch.int64)
    for c in mine:
        m = min(chunk, total - c * chunk)
        chars = synth.device_patterns(t_seq, m, length, seed=4000 + c).numpy()
        sp, ep, _ = ora.find_batch(chars, np.arange(m + 1, dtype=np.uint64) * np.uint64(length))
        counts += torch.tensor([m, int(np.count_nonzero(sp <= ep))])
    dist.all_reduce(counts)
    # configs[4]: one batch of mixed lengths, dealt in length order
    all_chars, all_offsets = synth.device_mixed_length_patterns(t_seq, sites, alt, 801, 16, 120, seed=5, error_rate=0.02)
    my_chars, my_offsets, ids = synth.device_shard_by_length(all_chars, all_offsets, rank, world)
    offs, vals, _ = orc.mem_batch(ora, olcp, my_chars.numpy(), my_offsets.numpy().astype(np.uint64))
    mem = torch.tensor([int(ids.numel()), int(offs[-1])], dtype=torch.int64)
    dist.all_reduce(mem)
    if rank == 0:
        whole_offs, _, _ = orc.mem_batch(ora, olcp, all_chars.numpy(), all_offsets.numpy().astype(np.uint64))
        np.save(os.path.join(out_dir, "strong.npy"), np.array([int(counts[0]), int(counts[1]), int(mem[0]), int(mem[1]), int(whole_offs[-1])]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_strong_scaling_legs(tmp_path):
    port = free_port()
    mp.spawn(strong_scaling_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    queries, found, patterns, matches, whole_matches = np.load(os.path.join(str(tmp_path), "strong.npy"))
    assert queries == 6500 and found == 6500                          # every chunk once, everything sampled from the text occurs
    assert patterns == 801 and matches == whole_matches and matches > 801
