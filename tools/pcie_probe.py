#!/usr/bin/env python
"""Host <-> device copy ceiling of this box, ALONE and with N GPUs copying CONCURRENTLY (the ceiling of bench.py's
end-to-end number: every frame moves 1.1 GB each way between pinned host memory and its GPU).

    python tools/pcie_probe.py                 # one GPU: H2D alone, D2H alone, full duplex
    python tools/pcie_probe.py --gpus 1,2,4,8  # N processes (one per GPU), barrier-synchronised; aggregate per direction

For every N it prints one JSON line: per-GPU and aggregate GB/s for H2D only, D2H only and full duplex (both directions
at once, which is what the streaming pipeline does), with the copies issued as one monolithic 1.1 GB `copy_` per
direction (what cra5_b200.stream.FramePipeline does) and as 8 chunks on two streams per direction. Also prints what the
platform says about NUMA placement (sysfs) and the CPU count, because pinned pages live on the node of the allocating
thread. Committed output: profiles/r2_pcie_probe.txt.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N_ELEMS = 268 * 721 * 1440          # one ERA5 frame, fp32: 1.113 GB


def worker(rank, world, barrier, q, reps):
    import torch
    torch.cuda.set_device(rank)
    h_in = torch.empty(N_ELEMS).pin_memory()
    h_out = torch.empty(N_ELEMS).pin_memory()
    d_in = torch.empty(N_ELEMS, device="cuda")
    d_out = torch.randn(N_ELEMS, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(4)]
    chunk = (N_ELEMS + 7) // 8

    def issue(h2d, d2h, chunked):
        if not chunked:
            if h2d:
                with torch.cuda.stream(streams[0]):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(streams[1]):
                    h_out.copy_(d_out, non_blocking=True)
            return
        for c in range(8):
            a, b = c * chunk, min(N_ELEMS, (c + 1) * chunk)
            if h2d:
                with torch.cuda.stream(streams[c % 2]):
                    d_in[a:b].copy_(h_in[a:b], non_blocking=True)
            if d2h:
                with torch.cuda.stream(streams[2 + c % 2]):
                    h_out[a:b].copy_(d_out[a:b], non_blocking=True)

    def run(h2d, d2h, chunked):
        issue(h2d, d2h, chunked)            # warm
        torch.cuda.synchronize()
        barrier.wait()
        t = time.perf_counter()
        for _ in range(reps):
            issue(h2d, d2h, chunked)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        barrier.wait()
        return N_ELEMS * 4 * reps / dt / 1e9

    out = {}
    for name, (h, d, c) in (("h2d", (1, 0, 0)), ("d2h", (0, 1, 0)), ("duplex", (1, 1, 0)), ("duplex_chunked", (1, 1, 1))):
        out[name] = run(h, d, c)
    q.put((rank, out))


def topology():
    info = {"cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    try:
        import torch
        from cra5_b200.stream import cuda_pci_bus_id, gpu_numa_cpus
        info["gpus"] = torch.cuda.device_count()
        nodes = []
        for i in range(torch.cuda.device_count()):
            bus = cuda_pci_bus_id(i)
            try:
                node = open(f"/sys/bus/pci/devices/{bus.lower()}/numa_node").read().strip()
            except OSError:
                node = "?"
            cpus = gpu_numa_cpus(bus)
            nodes.append({"gpu": i, "bus": bus, "numa_node": node, "node_cpus": len(cpus) if cpus else None})
        info["gpu_numa"] = nodes
        info["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")) \
            if os.path.isdir("/sys/devices/system/node") else None
    except Exception as e:
        info["error"] = repr(e)
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--reps", type=int, default=6)
    a = ap.parse_args()
    import torch
    import torch.multiprocessing as mp
    mp.set_start_method("spawn", force=True)
    print(json.dumps({"topology": topology()}), flush=True)
    have = torch.cuda.device_count()
    for n in [int(v) for v in a.gpus.split(",")]:
        if n > have:
            print(json.dumps({"n_gpus": n, "skipped": f"only {have} GPUs visible"}), flush=True)
            continue
        barrier, q = mp.Barrier(n), mp.Queue()
        procs = [mp.Process(target=worker, args=(r, n, barrier, q, a.reps)) for r in range(n)]
        for p in procs:
            p.start()
        res = dict(q.get(timeout=600) for _ in range(n))
        for p in procs:
            p.join()
        line = {"n_gpus": n}
        for k in ("h2d", "d2h", "duplex", "duplex_chunked"):
            per = [res[r][k] for r in range(n)]
            line[k] = {"per_gpu_min": round(min(per), 1), "per_gpu_max": round(max(per), 1),
                       "aggregate_per_direction": round(sum(per), 1)}
        line["frames_per_s_ceiling_duplex"] = round(line["duplex"]["aggregate_per_direction"] / (N_ELEMS * 4 / 1e9), 1)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
