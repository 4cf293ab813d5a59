"""Multi-GPU frame streaming: hourly ERA5 frames are independent units (reference: test.py:13 loops timestamps, no
temporal context in the model), so N GPUs run N replicas and frame i goes to rank i mod N. No collective touches the
data path; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to agree on timings and to gather
per-rank byte counts."""
from __future__ import annotations

from typing import Iterable, List, Sequence


def shard_frames(n_frames: int, rank: int, world: int) -> List[int]:
    """indices of the frames rank `rank` of `world` processes (rank-strided, like a DistributedSampler without padding)"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"invalid rank {rank} for world size {world}")
    return list(range(rank, n_frames, world))


def max_over_ranks(seconds: float, device=None) -> float:
    """wall/device time of the slowest rank (multi-GPU numbers are always the max over ranks)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(values: Sequence[int], device=None) -> List[List[int]]:
    """per-rank integer lists (e.g. compressed bytes per frame), padded with -1, gathered to every rank"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(values)]
    n = torch.tensor([len(values)], dtype=torch.int64, device=device)
    dist.all_reduce(n, op=dist.ReduceOp.MAX)
    buf = torch.full((int(n.item()),), -1, dtype=torch.int64, device=device)
    buf[: len(values)] = torch.tensor(list(values), dtype=torch.int64, device=device)
    out = [torch.empty_like(buf) for _ in range(dist.get_world_size())]
    dist.all_gather(out, buf)
    return [[int(v) for v in t.tolist() if v >= 0] for t in out]


def stream_frames(codec, frames: Iterable, rank: int, world: int, n_frames: int):
    """compress+decompress this rank's share of `frames` (an indexable of (C,H,W) tensors); yields (index, strings,
    x_hat). `codec` is a cra5_b200.vaeformer.VAEformer living on this rank's GPU."""
    for i in shard_frames(n_frames, rank, world):
        x = frames[i]
        out = codec.compress(x.unsqueeze(0) if x.dim() == 3 else x)
        rec = codec.decompress(out["strings"], out["z_shape"])
        yield i, out["strings"], rec["x_hat"]


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' (sysfs cpulist syntax) -> [0, 1, 2, 3, 8, 10, 11]"""
    cpus: List[int] = []
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_cpus(pci_bus_id: str, sysfs: str = "/sys"):
    """CPUs of the NUMA node a GPU hangs off ('0000:1b:00.0' -> [0..31]); None when the platform does not say (single
    node, virtualised PCI topology, sysfs absent)"""
    import os
    try:
        with open(os.path.join(sysfs, "bus", "pci", "devices", pci_bus_id.lower(), "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(os.path.join(sysfs, "devices", "system", "node", f"node{node}", "cpulist")) as f:
            cpus = parse_cpulist(f.read())
        return cpus or None
    except (OSError, ValueError):
        return None


def cuda_pci_bus_id(device) -> str:
    """sysfs-style PCI address of a CUDA device"""
    import torch
    pr = torch.cuda.get_device_properties(device)
    return f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"


class near_gpu:
    """Context manager: run the enclosed host allocations on the NUMA node next to `device` (additive; matters for the
    N-GPU streaming path, where every rank moves 2 x 1.1 GB per frame between pinned host memory and its GPU -- pinned
    pages are placed on the node of the thread that allocates them, and a buffer on the far socket halves the PCIe rate
    and loads the inter-socket link). Restores the previous CPU affinity on exit; a no-op when the topology is unknown.

        with near_gpu(dev) as bound:        # bound: True if the affinity was narrowed
            buf = torch.empty(shape).pin_memory()
    """

    def __init__(self, device=None, pci_bus_id: str = None, sysfs: str = "/sys"):
        self.device, self.bus, self.sysfs = device, pci_bus_id, sysfs
        self.prev = None

    def __enter__(self):
        import os
        try:
            bus = self.bus or cuda_pci_bus_id(self.device)
            cpus = gpu_numa_cpus(bus, self.sysfs)
            if not cpus:
                return False
            prev = os.sched_getaffinity(0)
            want = set(cpus) & set(prev)
            if not want or want == set(prev):
                return False
            os.sched_setaffinity(0, want)
            self.prev = prev
            return True
        except Exception:
            return False

    def __exit__(self, *exc):
        import os
        if self.prev is not None:
            try:
                os.sched_setaffinity(0, self.prev)
            finally:
                self.prev = None
        return False


class CodecLanes:
    """Several codec lanes on ONE GPU (experimental until timed on a B200; bench.py --lanes, tools/check_overlap.py).

    Lane k = its own library handle (`VAEformer.replica()`: own workspace and bitstream buffers, shared weights) on its
    own CUDA stream, driven by its own host thread; item i runs on lane i mod L. Frames are independent units, so the
    lanes never exchange data and every item's result is bit-identical to the single-lane result. What the lanes buy
    is occupancy: a persistent GEMM with 2.2 waves of tiles leaves most SMs idle during its last wave, and
    `latent_to_bin` synchronises the host on the bitstream length -- with a second lane queued, the block scheduler
    fills those holes with the other frame's kernels. The library is re-entrant across handles (thread-local error
    and profiler state, per-handle buffers); ctypes releases the GIL during each call, so the lanes' launches overlap.

        lanes = CodecLanes(net, lanes=2)
        results, ms = lanes.run(lambda codec, i: roundtrip(codec, frames[i % 2]), n_items=10, timed=True)
    """

    def __init__(self, net, lanes: int = 2, streams=None):
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        self.nets = [net] + [net.replica() for _ in range(lanes - 1)]
        self.device = getattr(net, "device", None)
        if streams is None:
            import torch
            # resolve "cuda" to an index on the caller's thread: a new host thread starts on device 0
            idx = self.device.index if getattr(self.device, "index", None) is not None else torch.cuda.current_device()
            self.device = torch.device("cuda", idx)
            streams = [torch.cuda.Stream(self.device) for _ in range(lanes)]
        if len(streams) != lanes:
            raise ValueError("one stream per lane")
        self.streams = streams

    @staticmethod
    def _enter(stream):
        import contextlib
        if stream is None:                 # CPU tests of the orchestration: no CUDA streams
            return contextlib.nullcontext()
        import torch
        return torch.cuda.stream(stream)

    def launches(self) -> int:
        """kernel launches counted by the library on the lanes' host threads during the last run()"""
        return int(sum(self._launches))

    def run(self, fn, n_items: int, timed: bool = False):
        """fn(codec, i) for i in range(n_items), item i on lane i mod L; returns the results in item order (and, with
        timed=True, the device time in ms between a start event every lane waits for and an end event recorded after
        every lane has finished -- CUDA events, not wall clock)."""
        import threading
        L = len(self.nets)
        results = [None] * n_items
        errors = [None] * L
        self._launches = [0] * L
        cuda = self.streams[0] is not None
        start = end = cur = None
        if cuda:
            import torch
            cur = torch.cuda.current_stream(self.device)
            if timed:
                start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                start.record(cur)
            for s in self.streams:
                s.wait_stream(cur)         # inputs produced on the caller's stream are visible to every lane

        def lane(k):
            try:
                if cuda:
                    import torch
                    torch.cuda.set_device(self.device)   # a new host thread starts on device 0
                with self._enter(self.streams[k]):
                    n0 = self._count()
                    for i in range(k, n_items, L):
                        results[i] = fn(self.nets[k], i)
                    self._launches[k] = self._count() - n0
            except BaseException as e:  # re-raised on the caller's thread
                errors[k] = e

        threads = [threading.Thread(target=lane, args=(k,), name=f"cra5-lane-{k}") for k in range(1, L)]
        for t in threads:
            t.start()
        lane(0)                            # lane 0 runs on the caller's thread
        for t in threads:
            t.join()
        if cuda:
            for s in self.streams:
                cur.wait_stream(s)         # results are ordered before whatever the caller enqueues next
        for e in errors:
            if e is not None:
                raise e
        if timed:
            if not cuda:
                return results, 0.0
            end.record(cur)
            end.synchronize()
            return results, start.elapsed_time(end)
        return results

    def _count(self) -> int:
        """the library's launch counter is per host thread"""
        if self.streams[0] is None:
            return 0
        import ctypes
        from cra5_b200 import _lib
        c = ctypes.c_uint64()
        _lib.check(_lib.lib.cra5_launch_count(ctypes.byref(c)))
        return int(c.value)


class FramePipeline:
    """Host-to-host streaming of frames through one GPU with the PCIe copies hidden behind compute.

    Three CUDA streams: H2D of frame i+1 and D2H of reconstruction i-1 run while frame i is being compressed /
    decompressed (PCIe is full duplex; a 268x721x1440 frame is 1.1 GB each way, ~21 ms at 53 GB/s, about the same as
    the GPU time of the codec itself). The reference loops synchronously over timestamps (test.py:13): read, `.to(device)`,
    compress, decompress, `.cpu()`.

        pipe = FramePipeline(api)                      # api: cra5_b200.api.cra5_api
        for idx, strings, x_hat_host in pipe.run(host_frames, out_buffers): ...

    `host_frames`: indexable of pinned (C, H, W) fp32 host tensors in physical units; `out_buffers`: >= 2 pinned host
    tensors that receive the (normalised) reconstructions round-robin -- an entry is valid until it is reused.
    """

    def __init__(self, api, roundtrip: bool = True, bin_dir: str = None):
        """bin_dir: when given, every frame's strings go through the reference's `.bin` container on disk -- written
        with api/utils.write_bin (cra5_api.py:108-116) and read back by `api.bin_to_latent(path)` (:127-144) -- instead
        of being handed to the decoder in memory; four files are reused round-robin."""
        import torch
        self.api = api
        self.roundtrip = roundtrip
        self.bin_dir = bin_dir
        self.dev = torch.device(api.device)
        self.copy_in = torch.cuda.Stream(self.dev)
        self.copy_out = torch.cuda.Stream(self.dev)

    def run(self, host_frames, out_buffers, n_frames=None):
        import torch
        api, dev = self.api, self.dev
        n = len(host_frames) if n_frames is None else n_frames
        if n == 0:
            return
        main = torch.cuda.current_stream(dev)
        shape = tuple(host_frames[0].shape)
        dev_in = [torch.empty(shape, device=dev, dtype=torch.float32) for _ in range(2)]
        in_ready = [torch.cuda.Event() for _ in range(2)]
        in_free = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            b = i % 2
            with torch.cuda.stream(self.copy_in):
                self.copy_in.wait_event(in_free[b])          # the compute that last read this buffer has finished
                dev_in[b].copy_(host_frames[i % len(host_frames)], non_blocking=True)
                in_ready[b].record(self.copy_in)

        for b in range(2):
            in_free[b].record(main)
        prefetch(0)
        if n > 1:
            prefetch(1)
        pending = None  # (index, strings, host buffer, event): reconstruction still in flight to the host
        for i in range(n):
            b = i % 2
            main.wait_event(in_ready[b])
            y = api.encode_to_latent(data=dev_in[b])          # fused normalise + g_a + quant_conv
            in_free[b].record(main)
            # frame i+2 reuses this buffer: queue its copy NOW, behind frame i+1's on the copy stream, so that the H2D
            # engine never waits for the host loop to come round again (queued at the top of the next iteration it sat
            # idle ~2 ms per frame: the loop is paced by the D2H of the previous reconstruction)
            if i + 2 < n:
                prefetch(i + 2)
            out = api.latent_to_bin(y)                        # h_a, h_s, quantise, rANS -> host bytes
            if not self.roundtrip:
                yield i, out["strings"], None
                continue
            if self.bin_dir is not None:
                from .api.utils import write_bin
                path = f"{self.bin_dir}/frame_{i % 4}.bin"
                write_bin(path, out["strings"], out["z_shape"])
                y_hat = api.bin_to_latent(path)
            else:
                y_hat = api.net.decompress(out["strings"], out["z_shape"], return_format="latent")
            x_hat = api.latent_to_reconstruction(y_hat)       # post_quant_conv + g_s
            done = torch.cuda.Event()
            done.record(main)
            ob = out_buffers[i % len(out_buffers)]
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(done)
                ob.copy_(x_hat[0], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_out)
            x_hat.record_stream(self.copy_out)
            if pending is not None:
                pending[3].synchronize()                      # reconstruction i-1 has landed in host memory
                yield pending[0], pending[1], pending[2]
            pending = (i, out["strings"], ob, ev)
        if pending is not None:
            pending[3].synchronize()
            yield pending[0], pending[1], pending[2]
