"""bench.py end to end on the host emulation of the engine (tests/emu): every leg of the bench line -- device-resident
find, the host entry point with its automatic packing policy, the random-pattern leg, the locate leg, the CPU
baseline and the roofline accounting -- runs at a small size without a GPU, so that a mistake in the script is found
here and not by the one run on the box.  torch.cuda is replaced by host stand-ins for the duration of the test;
the numbers printed are meaningless and are not looked at, the structure of the line is."""
import ctypes
import json
import sys
import time

import pytest
import torch


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, 1000.0 * (other.t - self.t))


class _Stream:
    cuda_stream = 0


def host_stand_ins(monkeypatch):
    """The emulated engine behind capi.lib() and torch.cuda replaced by host stand-ins (device memory is host memory)."""
    from emu import build_emu
    from gcsa2_b200 import capi
    emulated = capi._bind(ctypes.CDLL(build_emu.build()))
    monkeypatch.setattr(capi, "_lib", emulated)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: _Stream())
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    from helpers import as_emulated_device_memory as on_device
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: on_device(self.clone()))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    real_empty, real_tensor, real_empty_like = torch.empty, torch.tensor, torch.empty_like
    monkeypatch.setattr(torch, "empty", lambda *a, **k: (on_device if "device" in k else (lambda t: t))(
        real_empty(*a, **{x: y for x, y in k.items() if x != "device"})))
    monkeypatch.setattr(torch, "empty_like", lambda t, *a, **k: on_device(real_empty_like(t, *a, **k)))
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{x: y for x, y in k.items() if x != "device"}))
    # the device-side generators of configs[3] run on the host, their results stand for device memory
    from types import SimpleNamespace
    from gcsa2_b200 import synth
    real_sequence, real_patterns = synth.device_sequence, synth.device_patterns
    monkeypatch.setattr(synth, "device_sequence", lambda length, seed, **k: on_device(real_sequence(length, seed, device="cpu")))
    monkeypatch.setattr(synth, "device_patterns", lambda *a, **k: on_device(real_patterns(*a, **k)))
    real_mixed, real_shard = synth.device_mixed_length_patterns, synth.device_shard_by_length
    monkeypatch.setattr(synth, "device_mixed_length_patterns", lambda *a, **k: tuple(on_device(t) for t in real_mixed(*a, **k)))
    monkeypatch.setattr(synth, "device_shard_by_length", lambda *a, **k: tuple(on_device(t.clone()) for t in real_shard(*a, **k)))
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda d: SimpleNamespace(total_memory=192 << 30))
    return build_emu


def test_bench_line_on_the_emulated_engine(monkeypatch, capsys):
    import bench
    build_emu = host_stand_ins(monkeypatch)
    monkeypatch.setenv("GCSA_B200_HOST_PACK_THREADS", "2")
    monkeypatch.delenv("GCSA_B200_HOST_PACK", raising=False)         # the default: raw copies and packing share the batch
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "2", "--warmup", "1", "--ref-mbp", "0.2", "--queries", "1100000",
                                      "--kmer-table-k", "8", "--locate-mbp", "0.2", "--locate-queries", "40000", "--cpu-sample", "20000",
                                      "--cfg4-mbp", "0.05", "--cfg4-queries", "250000", "--cfg4-chunk", "100000", "--cfg4-steps", "1",
                                      "--mem-patterns", "3000", "--mem-steps", "1", "--mem-cpu-sample", "1000"])
    bench.main()
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    with open(build_emu.OUT + "/bench_line.json", "w") as f:         # for a look at the whole line (git-ignored)
        json.dump(line, f, indent=1)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "locate", "secondary"):
        assert key in line, key
    assert line["found"] == line["queries"] == 1_100_000
    assert line["config"]["index"]["fused_table"] is True
    pack = line["e2e"]["host_pack"]
    assert line["e2e"]["matches_device_leg"] and pack["policy"] == "auto"
    assert (pack["chunks"] == 9 and 0 <= pack["packed_chunks"] <= 7) or (pack["chunks"] == 5 and pack["packed_chunks"] == 0)   # shared, or raw only
    assert 8 * 1_100_000 <= line["e2e"]["h2d_bytes_per_step"] <= 32 * 1_100_000
    assert line["cpu_baseline"]["parity_on_sample"] and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["secondary"]["parity_on_sample"] and line["secondary"]["found"] < 1_100_000 // 2
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    c5 = line["cfg5"]
    assert "error" not in c5, c5
    assert c5["patterns"] == 3000 and c5["matches"] > 3000 and c5["cpu_baseline"]["parity_on_sample"] and "roofline" in c5
    c4 = line["cfg4"]
    assert "error" not in c4, c4
    assert c4["queries"] == c4["found"] == 250_000 and c4["scaling"] == "strong" and c4["cpu_baseline"]["parity_on_sample"]
    assert c4["config"]["index"]["path_nodes"] == 50_002 and "roofline" in c4
    loc = line["locate"]
    assert "error" not in loc, loc
    assert loc["positions"] >= 40_000 and loc["e2e"]["matches_device_leg"] and loc["cpu_baseline"]["parity_on_sample"]
    wide = loc["wide_ranges"]
    assert [w["pattern_length"] for w in wide] == [10, 8, 6] and all(w["cpu_baseline"]["parity_on_sample"] and w["positions"] > 0 for w in wide)
    assert wide[0]["path_nodes_per_range"] < wide[1]["path_nodes_per_range"] < wide[2]["path_nodes_per_range"]


def test_smoke_on_the_emulated_engine(monkeypatch, capsys):
    """__graft_entry__.smoke() -- the one call the driver makes on the box before the bench -- on the emulated engine."""
    import __graft_entry__ as entry
    from emu import build_emu
    from gcsa2_b200 import capi
    monkeypatch.setattr(capi, "_lib", capi._bind(ctypes.CDLL(build_emu.build())))
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out


def test_bench_ops_on_the_emulated_engine(monkeypatch, capsys, tmp_path):
    """scripts/bench_ops.py (every operation on the variation-graph config, next to the CPU oracle) at a small size."""
    import importlib.util
    import os
    host_stand_ins(monkeypatch)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_ops", os.path.join(root, "scripts", "bench_ops.py"))
    bench_ops = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench_ops)
    out = str(tmp_path / "ops.json")
    monkeypatch.setattr(sys, "argv", ["bench_ops.py", "--mbp", "0.1", "--queries", "20000", "--steps", "1", "--cpu-sample", "5000",
                                      "--kmer-table-k", "8", "--ops", "count,locate,parent,mem,kmers,compare", "--out", out])
    bench_ops.main()
    with open(out) as f:
        rows = json.load(f)
    names = " | ".join(r["op"] for r in rows)
    for expected in ("find", "count", "locate", "parent", "depth", "GCSA_B200_MEM_JUMP=0", "GCSA_B200_MEM_JUMP=1", "countKMers(k=12)", "compareKMers(k=16)"):
        assert expected in names, (expected, names)
    assert all(r["parity_on_sample"] for r in rows), [r["op"] for r in rows if not r["parity_on_sample"]]
