"""The device-side generators of the benchmark's configs[3] and configs[4] legs (gcsa2_b200/synth.py: counter-based, so that
every rank regenerates the same data) and the way those legs share a job between the ranks: run here on CPU tensors."""
import numpy as np
import torch

from gcsa2_b200 import dist, synth


def test_counter_sequence_is_the_same_on_host_and_device_and_uniform():
    host = synth.counter_sequence(200_000, seed=4)
    dev = synth.device_sequence(200_000, seed=4, device="cpu", chunk=1 << 16).numpy()
    assert (host == dev).all() and host.min() == 1 and host.max() == 4
    counts = np.bincount(host)[1:]
    assert abs(counts - 50_000).max() < 1500
    assert (synth.counter_sequence(1000, seed=5) != host[:1000]).any()


def test_device_patterns_are_substrings_of_the_reference():
    seq = torch.from_numpy(synth.counter_sequence(50_000, seed=4))
    n, length = 3000, 32
    chars = synth.device_patterns(seq, n, length, seed=4001, chunk=1024).numpy()
    starts = synth.device_pattern_starts(50_000, n, length, seed=4001, device="cpu").numpy()
    assert starts.min() >= 0 and starts.max() <= 50_000 - length
    text = synth.COMP2CHAR[seq.numpy()]
    for i in (0, 1, 999, 2999):
        assert (chars[i * length:(i + 1) * length] == text[starts[i]:starts[i] + length]).all()
    again = synth.device_patterns(seq, n, length, seed=4001).numpy()
    assert (again == chars).all()                                      # every rank regenerates the same chunk


def test_mixed_length_patterns_and_their_shards():
    seq = synth.random_sequence(100_000, seed=3)
    graph, sites, alt = synth.snp_graph(seq, seed=3, snp_rate=0.01)
    chars, offsets = synth.device_mixed_length_patterns(torch.from_numpy(seq), sites, alt, 4000, 16, 256, seed=5, error_rate=0.01)
    lengths = np.diff(offsets.numpy())
    assert lengths.min() >= 16 and lengths.max() <= 256 and 120 < lengths.mean() < 152
    assert set(np.unique(chars.numpy())) <= set(b"ACGT")
    # the shards of every world size: the numpy rule of dist.shard_patterns_by_length, a partition, the same mix of lengths
    for world in (1, 2, 3, 4, 8):
        seen = []
        for rank in range(world):
            c, o, ids = synth.device_shard_by_length(chars, offsets, rank, world)
            hc, ho, hids = dist.shard_patterns_by_length(chars.numpy(), offsets.numpy().astype(np.uint64), rank, world)
            assert (c.numpy() == hc).all() and (o.numpy().astype(np.uint64) == ho).all() and (ids.numpy() == hids).all()
            seen.append(ids.numpy())
            mine = np.diff(o.numpy())
            assert abs(mine.mean() - lengths.mean()) < 4 and abs(len(mine) - 4000 / world) <= 1
            for j in (0, len(mine) - 1):
                g = int(ids[j])
                assert (c[int(o[j]):int(o[j + 1])].numpy() == chars[int(offsets[g]):int(offsets[g + 1])].numpy()).all()
        assert (np.sort(np.concatenate(seen)) == np.arange(4000)).all()


def test_cfg4_chunks_are_dealt_once():
    """bench.py's configs[3] leg: chunk c of the job belongs to rank c % world -- every chunk exactly once for every N."""
    total, chunk = 1_000_000_000, 125_000_000
    n_chunks = (total + chunk - 1) // chunk
    for world in (1, 2, 4, 8, 3):
        owned = [c for rank in range(world) for c in range(n_chunks) if c % world == rank]
        assert sorted(owned) == list(range(n_chunks))
        sizes = [sum(min(chunk, total - c * chunk) for c in range(n_chunks) if c % world == rank) for rank in range(world)]
        assert sum(sizes) == total and (world not in (1, 2, 4, 8) or max(sizes) == min(sizes))
