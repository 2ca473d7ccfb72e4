#!/usr/bin/env python
"""bench.py — QPS @ 99% recall@10 of the diskann index-scan hot path on B200.

A "step" is one pass of the hot path (amrescan preparation -> StreamingDiskANN beam search over
SBQ codes -> exact f32 rerank window) over one batch of synthetic queries against a 1M x 768-d
index resident in HBM (BASELINE.json configs[1]).  Weak scaling: every rank holds a full replica
and its own batch; the only collective is the final all_gather of the top-k rows.

  python bench.py [--gpus N --steps K --warmup W]          our CUDA path (one JSON line)
  python bench.py --impl reference ...                     the CPU path (oracle port) on host cores

`value`  : whole-job queries/s with the query batch already in HBM (device-timed, max over ranks)
`e2e`    : the same metric through the host-buffer C-ABI call (dann_search_batch): pinned host
           queries in, (tid, dist) rows out, H2D/D2H inside the timed region
`roofline`: the beam-search kernel (dominant kernel) against the measured HBM peak
`cpu_baseline`: the CPU oracle timed on a bounded sample on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=1024, help="queries per GPU per step")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--data", default="lowrank", choices=["lowrank", "gaussian"])
    ap.add_argument("--fixture", default="vamana", choices=["vamana", "knn"],
                    help="graph of the synthetic index: the product's GPU batch Vamana builder over SBQ codes "
                         "(restates the reference build) or the exact-kNN + alpha-prune torch fixture")
    ap.add_argument("--bits", type=int, default=0, help="SBQ bits/dim (0 = reference default)")
    ap.add_argument("--L", type=int, default=0, help="fix search_list_size (0 = sweep for 99%% recall)")
    ap.add_argument("--rescore", type=int, default=0)
    ap.add_argument("--target-recall", type=float, default=0.99)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--sweep", action="store_true", help="print the whole recall/QPS sweep to stderr")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, for one),
# so fd 1 is pointed at stderr for the whole run and only emit() writes to the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# (search_list_size, rescore) in increasing cost (visits ~ L + rescore); recall is driven mostly by rescore
SWEEP = [(25, 50), (50, 50), (100, 50), (50, 100), (100, 100), (64, 150), (100, 150), (64, 200), (100, 200),
         (150, 200), (200, 200), (150, 250), (200, 250), (150, 300), (200, 300), (300, 300), (400, 400), (800, 400),
         (800, 800), (1000, 1000), (1600, 1000)]


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions (every ~5 ms through NVML;
    falls back to polling nvidia-smi when pynvml is unavailable)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.sm = []
        self.sm_max = None
        self.reasons = set()
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _sample_nvml(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                          ("hw_thermal_slowdown", 0x40), ("hw_power_brake_slowdown", 0x80)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            f = [x.strip() for x in out.split(",")]
            self.sm.append(float(f[0]))
            self.sm_max = float(f[1])
            for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                if f[3 + i].lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        self._stop.clear()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self._nvml else "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def build_fixture(args, device):
    """Dataset + index snapshot + queries; deterministic, identical on every rank."""
    import torch
    from tools import synth_index as si
    t0 = time.time()
    x = si.gen_dataset(args.n, args.dim, 0x5EED0010, args.data, device=device)
    if args.fixture == "vamana":
        snap, _, st = si.build_index_vamana(x, bits=args.bits or None, R=50, L_build=100, alpha=1.2, log=log)
    else:
        snap = si.build_index(x, bits=args.bits or None, R=50, log=log)
    log(f"[bench] index fixture ({args.fixture}): n={args.n} dim={args.dim} data={args.data} bits={snap.bits} "
        f"built in {time.time() - t0:.1f}s")
    return x, snap


def recall_at_k(tid, truth_nodes, snap, k):
    """tid [B,k] uint64 -> fraction of the exact top-k found."""
    t2n = None
    # synthetic heap tids are a bijection of node ids (make_heap_tids): node = block*2 + offset-1
    blk = (tid >> np.uint64(16)).astype(np.int64)
    off = (tid & np.uint64(0xFFFF)).astype(np.int64)
    nodes = blk * 2 + off - 1
    nodes[tid == np.uint64(0xFFFFFFFFFFFFFFFF)] = -1
    hits = 0
    for b in range(tid.shape[0]):
        hits += len(set(nodes[b].tolist()) & set(truth_nodes[b].tolist()))
    return hits / (tid.shape[0] * k)


def workload_name(args, bits):
    return (f"configs[1]: {args.n}x{args.dim}-d {args.data} ('Cohere-shape' synthetic) SBQ {bits}-bit "
            f"diskann index in HBM, batch={args.batch} queries/GPU/step, k={args.k}")


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for this path on the box's host cores.
    The Rust/pgrx extension cannot be built in this image (no rustc/cargo/Postgres), so this is the
    oracle port (oracle/oracle.cpp, the reference's AVX2+FMA flag family), one host thread per
    core, on the same snapshot, queries, k and operating point as our arm.  The GPU is used only to
    build the synthetic index fixture and the brute-force recall ground truth (untimed setup)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import oracle
    oracle.build_lib()
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    from tools import synth_index as si
    x, snap = build_fixture(args, device)
    B, k = args.batch, args.k
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or (B if cores >= 32 else max(64, 16 * cores))
    sample = min(sample, B)
    nb = args.warmup + args.steps
    q_first = si.gen_dataset(B, args.dim, 0x5EED0011, args.data, device=device)      # our arm's first rank-0 batch
    truth = si.ground_truth(x, q_first, k).cpu().numpy()
    q_first = q_first.cpu().numpy()
    del x
    torch.cuda.empty_cache()
    # same operating-point rule as our arm: first sweep point with recall@10 >= target
    points = [(args.L, args.rescore or 50)] if args.L else SWEEP
    chosen = None
    for (L, rescore) in points:
        tid, _, _, _ = oracle.scan_batch(snap, q_first, None, None, L, rescore, k, threads=0)
        rec = recall_at_k(tid, truth, snap, k)
        log(f"[bench/reference] sweep L={L} rescore={rescore}: recall@{k}={rec:.4f}")
        chosen = (L, rescore, rec)
        if rec >= args.target_recall:
            break
    L, rescore, recall = chosen
    rng = np.random.default_rng(0x5EED0012)
    qs = si.gen_dataset(nb * sample, args.dim, 0x5EED0013, args.data, device=device).cpu().numpy()
    for w in range(args.warmup):
        oracle.scan_batch(snap, qs[w * sample:(w + 1) * sample], None, None, L, rescore, k, threads=0)
    t0 = time.perf_counter()
    for s in range(args.warmup, nb):
        oracle.scan_batch(snap, qs[s * sample:(s + 1) * sample], None, None, L, rescore, k, threads=0)
    dt = time.perf_counter() - t0
    qps = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": "QPS @ 99% recall@10, 1Mx768-d SBQ diskann scan, k=10",
        "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 popcount + f32", "data": "synthetic",
        "config": {"workload": workload_name(args, snap.bits), "search_list_size": L, "rescore": rescore,
                   "recall_at_10": round(recall, 4),
                   "note": f"CPU arm: each step scans a bounded sample of {sample} queries of the workload"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {sample} queries, one host thread per core "
                                   "(oracle/oracle.cpp, -O2 -mavx2 -mfma -mpopcnt)"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pgvectorscale_b200 import diskann
    from pgvectorscale_b200.group import QueryShardGroup
    from tools import synth_index as si

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    B, k, dim = args.batch, args.k, args.dim

    x, snap = build_fixture(args, device)
    idx = diskann.DiskAnnIndex(snap, device=local_rank)
    log(f"[bench] rank {rank}: index in HBM: {idx.hbm_bytes / 1e9:.2f} GB")

    # ---- queries: per-rank distinct batches, (warmup+steps) of them so no step repeats a batch
    nb = args.warmup + args.steps
    # the first batch (recall / operating point / parity) is generated on its own so that the
    # reference arm can reproduce exactly the same queries
    q_all = torch.cat([si.gen_dataset(B, dim, 0x5EED0011 + 7919 * rank, args.data, device=device),
                       si.gen_dataset((nb - 1) * B, dim, 0x5EED0013 + 7919 * rank, args.data, device=device)])
    truth = si.ground_truth(x, q_all[:B], k).cpu().numpy()
    del x
    torch.cuda.empty_cache()

    stream = torch.cuda.current_stream(device)
    d_tid = torch.empty((B, k), dtype=torch.int64, device=device)
    d_dist = torch.empty((B, k), dtype=torch.float32, device=device)
    d_cnt = torch.empty(B, dtype=torch.int32, device=device)
    d_stats = torch.empty((B, 6), dtype=torch.int32, device=device)

    def run_device(qb, L, rescore):
        idx.search_batch_device(qb, k, L, rescore, d_tid, d_dist, d_cnt, d_stats, stream=stream.cuda_stream)

    # ---- operating point: first (L, rescore) of the sweep reaching the target recall
    points = [(args.L, args.rescore or 50)] if args.L else SWEEP
    chosen = None
    sweep_log = []
    for (L, rescore) in points:
        run_device(q_all[:B], L, rescore)          # first run of a plan pays for workspace (re)allocation
        run_device(q_all[:B], L, rescore)
        torch.cuda.synchronize(device)
        rec = recall_at_k(d_tid.cpu().numpy().view(np.uint64), truth, snap, k)
        t = idx.last_batch_timing()
        sweep_log.append({"L": L, "rescore": rescore, "recall": round(rec, 4),
                          "device_ms": round(t["total_ms"], 3), "qps_1gpu": round(B / t["total_ms"] * 1e3)})
        if args.sweep or rank == 0:
            log(f"[bench] sweep L={L} rescore={rescore}: recall@{k}={rec:.4f} device {t['total_ms']:.2f} ms "
                f"(search {t['search_ms']:.2f}, rerank {t['rerank_ms']:.2f}, retries {t['retries']})")
        if rec >= args.target_recall and chosen is None:
            chosen = (L, rescore, rec)
            if not args.sweep:
                break
    if chosen is None:
        L, rescore = points[-1]
        chosen = (L, rescore, sweep_log[-1]["recall"])
        log(f"[bench] WARNING: target recall {args.target_recall} not reached; reporting at L={L}")
    L, rescore, recall = chosen

    # ---- parity gate: identical TIDs vs the CPU oracle before any timing
    parity = None
    if not args.no_parity and rank == 0:
        from oracle import oracle
        oracle.build_lib()
        ns = min(64, B)
        run_device(q_all[:B], L, rescore)
        torch.cuda.synchronize(device)
        g_tid = d_tid[:ns].cpu().numpy().view(np.uint64)
        g_dist = d_dist[:ns].cpu().numpy()
        otid, odist, _, ostats = oracle.scan_batch(snap, q_all[:ns].cpu().numpy(), None, None, L, rescore, k, threads=0)
        same_ids = bool(np.array_equal(g_tid, otid))
        same_dist = bool(np.array_equal(g_dist.view(np.uint32), odist.view(np.uint32)))
        parity = {"queries": ns, "tids_identical": same_ids, "dist_bits_identical": same_dist}
        log(f"[bench] parity vs oracle on {ns} queries: ids {same_ids}, dist bits {same_dist}")
        if not same_ids:
            raise SystemExit("parity FAILED: returned row ids differ from the CPU oracle; refusing to report a number")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    shard = QueryShardGroup(None, k, device)

    # ---- device-resident timed region -----------------------------------------------------
    for w in range(args.warmup):
        run_device(q_all[w * B:(w + 1) * B], L, rescore)
        shard.gather_rows(d_tid, d_dist, B)
    launches0 = idx.kernel_launches
    search_ms = rerank_ms = prepare_ms = 0.0
    stat_sum = np.zeros(6, np.float64)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks = ClockSampler(local_rank)
    with clocks:
        ev0.record(stream)
        for s in range(args.warmup, nb):
            run_device(q_all[s * B:(s + 1) * B], L, rescore)
            shard.gather_rows(d_tid, d_dist, B)
            t = idx.last_batch_timing()
            search_ms += t["search_ms"]
            rerank_ms += t["rerank_ms"]
            prepare_ms += t["prepare_ms"]
        ev1.record(stream)
        barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = idx.kernel_launches - launches0
    # counters of the last batch -> algorithmic bytes: taken after timing
    st = d_stats.cpu().numpy().astype(np.float64)
    visits_q, dq_q = st[:, 0].mean(), st[:, 1].mean()

    # ---- end-to-end through the host-buffer C ABI ---------------------------------------------
    h_q = torch.empty((nb * B, dim), dtype=torch.float32).pin_memory()
    h_q.copy_(q_all)
    h_tid = torch.empty((B, k), dtype=torch.int64).pin_memory()
    h_dist = torch.empty((B, k), dtype=torch.float32).pin_memory()
    for w in range(args.warmup):
        idx.search_batch_ptrs(h_q[w * B:(w + 1) * B].data_ptr(), B, k, L, rescore, h_tid.data_ptr(), h_dist.data_ptr())
    barrier()
    with clocks:
        t0 = time.perf_counter()
        for s in range(args.warmup, nb):
            idx.search_batch_ptrs(h_q[s * B:(s + 1) * B].data_ptr(), B, k, L, rescore, h_tid.data_ptr(), h_dist.data_ptr())
            if world > 1:
                shard.gather_rows(torch.from_numpy(h_tid.numpy()).to(device), torch.from_numpy(h_dist.numpy()).to(device), B)
        barrier()
        e2e_s = time.perf_counter() - t0

    # ---- max over ranks ---------------------------------------------------------------------
    tm = torch.tensor([dev_ms, e2e_s * 1e3, search_ms, rerank_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, search_ms_max, rerank_ms_max = tm.tolist()
    total_q = args.steps * B * world
    value = total_q / (dev_ms / 1e3)
    e2e_value = total_q / (e2e_ms / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (beam search) ----------------------------------------
    peak, peak_src = measured_peak()
    code_bytes = idx.code_stride * 8
    alg_bytes_q = dq_q * code_bytes + visits_q * snap.R * 4           # SURVEY §8d per-query search bytes
    alg_bytes_launch = alg_bytes_q * B
    search_avg_ms = search_ms / args.steps
    achieved = alg_bytes_launch / (search_avg_ms / 1e3) / 1e9
    rerank_bytes_launch = B * (rescore + k - 1 if rescore else 0) * dim * 4
    rerank_avg_ms = rerank_ms / args.steps
    # DRAM traffic of one search-kernel launch from the committed `ncu --set full` capture of the same
    # operating point (profiles/r01_traffic.json); null when the capture is for another configuration
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if (tj.get("L"), tj.get("rescore"), tj.get("n"), tj.get("batch")) == (L, rescore, args.n, B):
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"kernel": "dann_search2_kernel<Ent32x21,3> (two warps per query)" if snap.R <= 64 else "dann_search_kernel", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": int(alg_bytes_launch), "avg_launch_ms": round(search_avg_ms, 4),
                "per_query": {"visits": round(visits_q, 1), "d_quantized": round(dq_q, 1), "code_bytes": code_bytes}}
    others = {"dann_rerank_kernel": {"alg_bytes_per_launch": int(rerank_bytes_launch),
                                     "avg_launch_ms": round(rerank_avg_ms, 4),
                                     "achieved_gbs": round(rerank_bytes_launch / max(rerank_avg_ms, 1e-9) / 1e6, 1)},
              "dann_prepare_kernel": {"avg_launch_ms": round(prepare_ms / args.steps, 4)}}

    # ---- stand-alone SBQ-distance kernel (the metric's named kernel) ---------------------------
    try:
        npairs = 64 * 1024 * 1024 if args.n >= 500_000 else 4 * 1024 * 1024
        g = torch.Generator(device=device)
        g.manual_seed(1)
        pn = torch.randint(0, snap.n, (npairs,), generator=g, device=device, dtype=torch.int32)
        pq = torch.randint(0, B, (npairs,), generator=g, device=device, dtype=torch.int32)
        qc = torch.empty((B, idx.code_stride), dtype=torch.int64, device=device)
        idx.prepare_queries(q_all[:B].contiguous(), None, qc)
        out = torch.empty(npairs, dtype=torch.int32, device=device)
        for _ in range(3):
            idx.sbq_distance(qc, pq, pn, out, stream=stream.cuda_stream)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        reps = 5
        a.record(stream)
        for _ in range(reps):
            idx.sbq_distance(qc, pq, pn, out, stream=stream.cuda_stream)
        b.record(stream)
        torch.cuda.synchronize(device)
        ms = a.elapsed_time(b) / reps
        sb = npairs * (code_bytes + 12)
        others["dann_sbq_distance_kernel"] = {"npairs": npairs, "alg_bytes_per_launch": sb, "avg_launch_ms": round(ms, 4),
                                              "achieved_gbs": round(sb / ms / 1e6, 1), "frac": round(sb / ms / 1e6 / peak, 4)}
        del pn, pq, out
    except Exception as e:  # the headline must not die on the side measurement
        others["dann_sbq_distance_kernel"] = {"error": str(e)}

    # ---- CPU baseline: the oracle port on this box's cores, bounded sample ------------------------
    from oracle import oracle
    oracle.build_lib()
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or max(64, min(B, 16 * cores))
    qs = q_all[:sample].cpu().numpy()
    oracle.scan_batch(snap, qs[: max(8, sample // 8)], None, None, L, rescore, k, threads=0)   # warm the page cache
    t0 = time.perf_counter()
    reps = 0
    while True:
        oracle.scan_batch(snap, qs, None, None, L, rescore, k, threads=0)
        reps += 1
        if time.perf_counter() - t0 > 8.0 or reps >= 20:
            break
    cpu_qps = reps * sample / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    n1 = max(8, min(64, sample))
    oracle.scan_batch(snap, qs[:n1], None, None, L, rescore, k, threads=1)
    cpu1_qps = n1 / (time.perf_counter() - t0)

    line = {
        "metric": "QPS @ 99% recall@10, 1Mx768-d SBQ diskann scan, k=10",
        "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 popcount + f32", "data": "synthetic",
        "config": {"workload": workload_name(args, snap.bits),
                   "search_list_size": L, "rescore": rescore, "recall_at_10": round(recall, 4),
                   "parallelism": f"query-shard x{world} (replicated index, all_gather of top-k)",
                   "l2_policy": f"index {idx.hbm_bytes / 1e9:.2f} GB >> 126 MB L2, random gathers, distinct queries every step",
                   "search_kernel": ("dann_search2_kernel<HV=1> DANN_HV_FLAGS=%s (alternative engine, see DESIGN.md)"
                                     % os.environ.get("DANN_HV_FLAGS", "all") if os.environ.get("DANN_HEAP_V2") == "1"
                                     else "dann_search2_kernel<HV=0> (the round-1 measured kernel)"),
                   "index_fixture": ("dann_build_graph: GPU batch Vamana over SBQ codes (R=50, L_build=100, alpha=1.2), "
                                     "the reference's build algorithm with batched insertion" if args.fixture == "vamana"
                                     else "tools/synth_index.py exact-kNN + alpha-prune torch fixture")},
        "recall_sweep": sweep_log,
        "parity": parity,
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": B * dim * 4,
                "d2h_bytes_per_step": B * k * 12, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernels": others,
        "cpu_baseline": {"value": cpu_qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{reps} x {sample} queries of the same workload, one host thread per core "
                                   f"(oracle/oracle.cpp, -O2 -mavx2 -mfma -mpopcnt); single-thread: {cpu1_qps:.0f} q/s"},
        "clocks": clocks.summary(),
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
