"""CPU tests of the host-side logic: geometry/config, state-dict schema, .bin container, CDF-table construction through
the C ABI, ABI symbol export, error behaviour without a GPU, and the N>1 sharding/timing plumbing on gloo."""
import ctypes
import io
import os
import re
import struct
import sys

import numpy as np
import pytest
import torch

from cra5_b200 import config as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shipped_geometry_matches_survey():
    cfg = C.cra5_268()
    assert cfg.grid == (72, 144) and cfg.tokens == 10368 and cfg.hyper_grid == (18, 36) and cfg.hyper_tokens == 648
    assert cfg.enc_blocks == 13 and cfg.dec_blocks == 12 and cfg.conv_head
    W = [(24, 24), (12, 48), (48, 12), None]
    assert cfg.enc_block_windows() == W * 3 + [None]          # W,W,W,G x3 + extra global
    assert cfg.dec_block_windows() == W * 3
    shapes = C.param_shapes(cfg)
    assert shapes["g_a.patch_embed.proj.weight"] == (1024, 268, 11, 10)
    assert shapes["g_s.final.weight"] == (1024, 268, 11, 10)
    assert shapes["quant_conv.weight"] == (512, 2048, 1, 1) and shapes["post_quant_conv.weight"] == (1024, 256, 1, 1)
    assert shapes["h_s.final.weight"] == (8192, 360) and shapes["h_a.quan_mlp.fc1.weight"] == (256, 360)
    n_params = sum(int(np.prod(s)) for s in shapes.values())
    assert abs(n_params / 1e6 - 404.68) < 0.05                  # SURVEY section 3.1 [probed]
    assert C.variant(159).in_chans == 159 and C.tiny_fullres().conv_head and not C.small_lowres().conv_head


def test_unsupported_geometry_is_rejected():
    with pytest.raises(ValueError):
        C.VaeformerConfig(patch_size=(11, 12), patch_stride=(10, 10)).validate()
    with pytest.raises(ValueError):
        C.VaeformerConfig(patch_size=(25, 10), patch_stride=(10, 10)).validate()
    with pytest.raises(ValueError):
        C.VaeformerConfig(num_heads=7).validate()


def test_bin_container_wire_format(tmp_path):
    from cra5_b200.api.utils import read_bin, write_bin
    strings = [[b"yyyy-stream"], [b"zz"]]
    p = tmp_path / "2024" / "frame.bin"
    n = write_bin(p, strings, (18, 36))
    raw = p.read_bytes()
    assert n == len(raw)
    # >I z_h, >I z_w, >I n_strings, then [>I len, bytes] per string (cra5_api.py:108-116)
    assert raw[:12] == struct.pack(">3I", 18, 36, 2)
    assert raw[12:16] == struct.pack(">I", 11) and raw[16:27] == b"yyyy-stream"
    assert raw[27:31] == struct.pack(">I", 2) and raw[31:] == b"zz"
    got, shape = read_bin(p)
    assert got == strings and tuple(shape) == (18, 36)
    with pytest.raises(ValueError):
        (tmp_path / "t.bin").write_bytes(raw[:20])
        read_bin(tmp_path / "t.bin")


def test_library_exports_every_declared_symbol():
    from cra5_b200 import _lib
    header = open(os.path.join(ROOT, "include", "cra5_b200.h")).read()
    names = re.findall(r"CRA5_API\s+[\w\s\*]+?\b(cra5_\w+)\s*\(", header)
    assert len(names) >= 20
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/cra5_b200.h but not exported"
    assert _lib.lib.cra5_abi_version() >= 1
    # every build variant (cra5_b200/build.py) exports the same ABI
    from cra5_b200 import build as B
    for variant in B.VARIANTS:
        path = B.lib_path(variant)
        assert os.path.exists(path), f"{path} missing: run python -m cra5_b200.build"
        v = ctypes.CDLL(path)
        for n_ in names:
            assert hasattr(v, n_), f"{n_} not exported by {os.path.basename(path)}"


def test_cdf_tables_through_c_abi_match_oracle():
    """update()'s integer tables (product path: torch pmf + cra5_pmf_to_quantized_cdf) == oracle == reference fixture"""
    from cra5_b200 import entropy_tables as ET
    from oracle import entropy_oracle as EO, weights
    assert ET.pmf_to_quantized_cdf(torch.tensor([0.1, 0.2, 0.7])).tolist() == [0, 6554, 19661, 65536]
    with pytest.raises(ValueError):
        ET.pmf_to_quantized_cdf(torch.tensor([0.5, float("nan")]))
    with pytest.raises(ValueError):
        ET.pmf_to_quantized_cdf(torch.tensor([0.0, 0.0]))
    g = ET.gaussian_conditional_tables(ET.get_scale_table())
    o = EO.gaussian_conditional_tables()
    assert torch.equal(g.quantized_cdf, o.cdf) and torch.equal(g.cdf_length, o.cdf_length) and torch.equal(g.offset, o.offset)
    cfg = C.tiny_fullres(69)
    sd = weights.seeded_state_dict(C.param_shapes(cfg), 7)
    e, eo = ET.entropy_bottleneck_tables(sd), EO.entropy_bottleneck_tables(sd)
    assert torch.equal(e.quantized_cdf, eo.cdf) and torch.equal(e.cdf_length, eo.cdf_length) and torch.equal(e.offset, eo.offset)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "tiny69.npz"))
    assert np.array_equal(e.quantized_cdf.numpy(), gold["eb_cdf"])
    # zero-probability bins get a count stolen from the narrowest donor
    cdf = ET.pmf_to_quantized_cdf(torch.tensor([0.5, 0.0, 0.5 - 1e-7, 1e-7])).tolist()
    assert all(b > a for a, b in zip(cdf, cdf[1:])) and cdf[-1] == 65536


def test_no_cpu_fallback_and_reference_errors():
    from cra5_b200 import zoo
    with pytest.raises(ValueError, match="Invalid metric"):      # zoo/image.py:316
        zoo.vaeformer_pretrained(268, metric="psnr")
    with pytest.raises(ValueError, match="Invalid quality"):     # zoo/image.py:319
        zoo.vaeformer_pretrained(0)
    with pytest.raises(ValueError, match="Invalid quality"):     # zoo/image.py:281-282
        zoo.vaeformer_pretrained(5)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU execution path"):
            zoo.vaeformer_pretrained(268)


def test_channel_table_and_api_bookkeeping():
    import json
    from cra5_b200.api import era5_268v as V
    names = V.channel_names()
    assert len(names) == 268 and names[0] == "z_1000" and names[37] == "q_1000" and names[74] == "u_1000" and names[-1] == "msl"
    stats = json.load(open(os.path.join(ROOT, "cra5_b200", "api", "era5_268v_stats.json")))["channels"]
    assert [r["name"] for r in stats] == names and all(r["std"] > 0 for r in stats)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from cra5_b200 import stream
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = stream.shard_frames(11, rank, world)
    slowest = stream.max_over_ranks(1.0 + rank)
    counts = stream.gather_counts([100 + i for i in mine])
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, slowest, counts))


def test_frame_sharding_world_size_2_gloo():
    """N > 1 path: rank-strided frames, no data-path collective, max-over-ranks timing"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, s0, c0), (r1, f1, s1, c1) = res
    assert f0 == [0, 2, 4, 6, 8, 10] and f1 == [1, 3, 5, 7, 9]
    assert sorted(f0 + f1) == list(range(11))
    assert s0 == s1 == 2.0
    assert c0 == c1 == [[100 + i for i in f0], [100 + i for i in f1]]
    from cra5_b200 import stream
    assert stream.shard_frames(5, 0, 1) == [0, 1, 2, 3, 4] and stream.max_over_ranks(3.5) == 3.5
    with pytest.raises(ValueError):
        stream.shard_frames(5, 2, 2)


class _FakeNvml:
    """stands in for pynvml: a clock that sags under load and a power-cap flag on the later samples"""
    NVML_CLOCK_SM = 1

    def __init__(self):
        self.n = 0
        self.bus = None

    def nvmlInit(self):
        pass

    def nvmlDeviceGetHandleByPciBusId(self, bus):
        self.bus = bus
        return "h"

    def nvmlDeviceGetHandleByIndex(self, i):
        return "h"

    def nvmlDeviceGetMaxClockInfo(self, h, kind):
        return 1965

    def nvmlDeviceGetClockInfo(self, h, kind):
        self.n += 1
        return 1965 if self.n < 3 else 1400

    def nvmlDeviceGetCurrentClocksEventReasons(self, h):
        return 0x4 if self.n > 4 else 0

    def nvmlDeviceGetPowerUsage(self, h):
        return 987000


def test_bench_clock_sampler_and_reference_arm_ranks(monkeypatch, capsys):
    """bench.py host logic: the clocks line takes many samples inside a short timed region (NVML thread), reports the
    median under load / the maximum clock / the union of slow-down reasons, degrades to 'no samples' without a driver;
    the reference arm prints nothing on ranks other than 0."""
    import time
    import types
    import bench
    nv = _FakeNvml()
    c = bench.ClockSampler(0, pci_bus_id="00000000:1B:00.0", period_s=0.005, nvml=nv)
    c.start()
    time.sleep(0.15)
    r = c.stop()
    assert nv.bus == b"00000000:1B:00.0"
    assert r["samples"] >= 10 and r["sm_mhz"] == 1400.0 and r["sm_max_mhz"] == 1965.0
    assert r["reasons"] == ["sw_power_cap"] and r["power_w_max"] == 987.0 and r["source"].startswith("nvml")
    if not torch.cuda.is_available():
        r = bench.ClockSampler(0).stop()                     # never started, no driver: still a well-formed dict
        assert r["sm_mhz"] is None and r["reasons"]
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    bench.run_reference(types.SimpleNamespace(channels=268, steps=1, warmup=1, gpus=2))
    assert capsys.readouterr().out == ""


class _FakeCodec:
    def __init__(self, tag="lane0"):
        self.tag = tag
        self.device = None

    def replica(self):
        return _FakeCodec("lane%d" % (int(self.tag[4:]) + 1))


def test_codec_lanes_orchestration():
    """CodecLanes host logic (no CUDA: streams=None): item i runs on lane i mod L, results come back in item order, the
    lanes run concurrently on their own host threads, a failing item's exception reaches the caller."""
    import threading
    import time
    from cra5_b200.stream import CodecLanes
    lanes = CodecLanes(_FakeCodec(), lanes=2, streams=[None, None])
    assert [n.tag for n in lanes.nets] == ["lane0", "lane0".replace("0", "1")]
    seen = []
    both = threading.Barrier(2, timeout=20)

    def work(codec, i):
        if i < 2:
            both.wait()                     # items 0 and 1 must be in flight at the same time (two threads)
        seen.append((codec.tag, i, threading.current_thread().name))
        time.sleep(0.001)
        return i * i

    res, ms = lanes.run(work, 7, timed=True)
    assert res == [i * i for i in range(7)] and ms == 0.0
    assert all(tag == f"lane{i % 2}" for tag, i, _ in seen)
    assert {name for tag, _, name in seen if tag == "lane1"} == {"cra5-lane-1"}
    assert lanes.run(lambda c, i: c.tag, 3) == ["lane0", "lane1", "lane0"]
    assert lanes.launches() == 0

    def boom(codec, i):
        if i == 3:
            raise ValueError("bad frame 3")
        return i

    with pytest.raises(ValueError, match="bad frame 3"):
        lanes.run(boom, 6)
    assert CodecLanes(_FakeCodec(), lanes=1, streams=[None]).run(lambda c, i: i + 1, 3) == [1, 2, 3]
    with pytest.raises(ValueError):
        CodecLanes(_FakeCodec(), lanes=0, streams=[])


def test_numa_local_pinned_allocation_helper(tmp_path):
    """near_gpu(): host buffers of the streaming path are allocated on the NUMA node next to the rank's GPU (sysfs
    topology faked here); affinity is restored afterwards; unknown topologies are a no-op."""
    from cra5_b200 import stream
    assert stream.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11] and stream.parse_cpulist("") == []
    bus = "0000:1b:00.0"
    d = tmp_path / "bus" / "pci" / "devices" / bus
    d.mkdir(parents=True)
    prev = os.sched_getaffinity(0)
    some = sorted(prev)[: max(1, len(prev) // 2)]
    (d / "numa_node").write_text("1\n")
    nd = tmp_path / "devices" / "system" / "node" / "node1"
    nd.mkdir(parents=True)
    (nd / "cpulist").write_text(",".join(str(c) for c in some) + "\n")
    assert stream.gpu_numa_cpus("0000:1B:00.0", sysfs=str(tmp_path)) == some
    with stream.near_gpu(pci_bus_id=bus, sysfs=str(tmp_path)) as bound:
        inside = os.sched_getaffinity(0)
    assert os.sched_getaffinity(0) == prev
    if len(prev) > 1:
        assert bound and inside == set(some)
    (d / "numa_node").write_text("-1\n")                       # platform does not say: leave the affinity alone
    with stream.near_gpu(pci_bus_id=bus, sysfs=str(tmp_path)) as bound:
        assert not bound and os.sched_getaffinity(0) == prev
    with stream.near_gpu(pci_bus_id="0000:ff:00.0", sysfs=str(tmp_path)) as bound:
        assert not bound


def test_reference_arm_times_whole_frames(monkeypatch, capsys):
    """`bench.py --impl reference`: a step is one WHOLE frame through compress + decompress (nothing extrapolated), the
    line reports the frames actually timed, and ms_per_step * steps is the wall time of the timed region."""
    import json
    import time
    import types
    import bench
    import oracle.cpu_baseline as CB

    calls = []

    class FakeTimer:
        kind = "port"

        def __init__(self, cfg, threads=None, seed=1234):
            self.cfg = cfg

        def frame(self):
            calls.append(time.perf_counter())
            time.sleep(0.02)
            return dict(encode_s=0.012, decode_s=0.008, total_s=0.02, bytes=1234)

        def describe(self, n):
            return f"{n} whole frames"

    monkeypatch.setattr(CB, "FrameTimer", FakeTimer)
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "1")
    t0 = time.perf_counter()
    bench.run_reference(types.SimpleNamespace(channels=268, steps=5, warmup=3, gpus=1))
    wall = time.perf_counter() - t0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 5 and line["warmup"] == 1
    assert len(calls) == 6                                            # 1 warm-up + 5 timed whole frames
    assert line["ms_per_step"] * line["steps"] / 1e3 <= wall          # fits the driver's clock
    assert abs(line["value"] - 1e3 / line["ms_per_step"]) < 1e-6
    assert line["e2e"]["value"] == line["value"] and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["sample"] == "5 whole frames"
    # the budget bounds the number of timed frames, never below 2
    monkeypatch.setenv("CRA5_REF_BUDGET_S", "0.05")
    bench.run_reference(types.SimpleNamespace(channels=268, steps=20, warmup=3, gpus=1))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["steps"] == 2 and line["steps_requested"] == 20


def test_whole_frame_timer_runs_the_real_codec():
    """oracle/cpu_baseline.FrameTimer on the reduced-width full-resolution geometry: a real compress + decompress"""
    from cra5_b200 import config as C
    from oracle.cpu_baseline import FrameTimer
    ft = FrameTimer(C.tiny_fullres(69), threads=4)
    r = ft.frame()
    assert r["total_s"] > 0 and abs(r["encode_s"] + r["decode_s"] - r["total_s"]) < 1e-3 and r["bytes"] > 1000
    assert "nothing extrapolated" in ft.describe(1)


def test_config_inferred_from_published_checkpoint_schema():
    """from_state_dict reads the geometry off the checkpoint (vaeformer.py:172 reads in_chans off the patch-embed
    weight); shapes only, so meta tensors stand in for the 1.6 GB of parameters"""
    from cra5_b200 import config as C
    for chans in (268, 159, 69):
        want = C.variant(chans) if chans != 268 else C.cra5_268()
        sd = {k: torch.empty(shape, device="meta") for k, shape in C.param_shapes(want).items()}
        got = C.config_from_state_dict(sd)
        assert got.to_dict() == want.to_dict()
    tiny = {k: torch.empty(s, device="meta") for k, s in C.param_shapes(C.tiny_fullres(69)).items()}
    with pytest.raises(ValueError, match="non-standard width"):
        C.config_from_state_dict(tiny)
