#!/usr/bin/env python
"""Names the limiter of bench.py's `e2e` at N > 1: per-rank pinned host <-> device bandwidth with
1 .. N ranks copying at the same time.  Plain cudaMemcpyAsync of one 508 MB buffer (the size of one pair's fp32
lookup results), one copy per buffer, CUDA events.

    python tools/d2h_bw.py                                        # 1 rank
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/d2h_bw.py

Every rank times: D2H alone, H2D alone, both directions at once -- first with all ranks active together, then
(world > 1) one rank at a time while the others idle.  Rank 0 prints one JSON line with per-rank GB/s, the
aggregate, the CPU / NUMA picture (os.sched_getaffinity, nvidia-smi topo) and how long pinning the buffer took."""
import json, os, subprocess, sys, time
import torch

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)

NBYTES = 12 * 324 * 136 * 240 * 4           # 507.6 MB: one pair's lookup results
t0 = time.perf_counter()
host = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
host2 = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
pin_s = time.perf_counter() - t0
d_a = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
d_b = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()


def gbs(fn, reps=5):
    fn(); torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s1.synchronize(); s2.synchronize()
    e1.record(); torch.cuda.synchronize()
    return reps * NBYTES / (e0.elapsed_time(e1) * 1e-3) / 1e9


def d2h():
    with torch.cuda.stream(s1):
        host.copy_(d_a, non_blocking=True)


def h2d():
    with torch.cuda.stream(s2):
        d_b.copy_(host2, non_blocking=True)


def both():
    d2h(); h2d()


def wait_streams(fn):
    def run():
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        fn()
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    return run


res = {"together": {"d2h": gbs(wait_streams(d2h)), "h2d": gbs(wait_streams(h2d)), "both_sum": 2 * gbs(wait_streams(both))}}
if world > 1:
    solo = {}
    # one rank at a time: time without the internal barrier
    def gbs_solo(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return reps * NBYTES / (e0.elapsed_time(e1) * 1e-3) / 1e9
    for r in range(world):
        barrier()
        if r == rank:
            solo = {"d2h": gbs_solo(wait_streams(d2h)), "h2d": gbs_solo(wait_streams(h2d))}
        barrier()
    res["alone"] = solo
res["pin_seconds_1GB"] = pin_s
res["cpus"] = sorted(os.sched_getaffinity(0))
if dist is not None:
    allres = [None] * world
    dist.gather_object(res, allres if rank == 0 else None, dst=0)
else:
    allres = [res]
if rank == 0:
    topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
    numa = subprocess.run(["bash", "-c", "lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"], capture_output=True, text=True).stdout
    line = {"tool": "d2h_bw", "n_ranks": world, "bytes_per_copy": NBYTES,
            "per_rank_GBps_all_ranks_together": [{k: round(v, 1) for k, v in r["together"].items()} for r in allres],
            "aggregate_d2h_GBps_together": round(sum(r["together"]["d2h"] for r in allres), 1),
            "aggregate_h2d_GBps_together": round(sum(r["together"]["h2d"] for r in allres), 1),
            "per_rank_GBps_one_rank_at_a_time": [{k: round(v, 1) for k, v in r.get("alone", r["together"]).items() if v} for r in allres],
            "pin_seconds_per_GB": [round(r["pin_seconds_1GB"], 3) for r in allres],
            "cpu_affinity_sizes": [len(r["cpus"]) for r in allres], "lscpu": numa.strip().splitlines(),
            "nvidia_smi_topo": topo.strip().splitlines()[:14]}
    print(json.dumps(line), flush=True)
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
