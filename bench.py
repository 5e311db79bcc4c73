#!/usr/bin/env python
"""bench.py -- scan-to-map registration throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement (oracle)

Workload (BASELINE.json configs[1], "c2"): 64-beam x 2048-azimuth synthetic LiDAR scans
(~130k points) -> voxel-grid downsample, 0.2 m leaf -> point-to-plane Gauss-Newton ICP, 10
iterations, threshold 0.5, against a 5M-point synthetic map with analytic normals.
A step registers one batch of `--scans-per-step` scans (distinct poses); N ranks each take
their own batch against their own replica of the map (no collective on the data path: the
work shards by scan, "weak" scaling).

Printed keys: value = scans/s with the raw scans already resident in HBM (CUDA events on the
library's stream); e2e = the same through the host-buffer API (pinned host scans -> H2D ->
align -> D2H results inside the timed region); roofline = the NN-search kernel's algorithmic
bytes / its CUDA-event time against the measured HBM copy bandwidth; cpu_baseline = the CPU
oracle on a bounded sample of the same workload on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-sensor-fusion_b200"))

WORKLOADS = {
    # name: (map points, beams, azimuths, leaf, mode, max_range)
    "c2": dict(map_points=5_000_000, beams=64, azimuths=2048, leaf=0.2, mode="gn_p2plane", max_range=100.0,
               scans_per_step=256,
               desc="64-beam scan (~130k pts) voxel 0.2 m + point-to-plane GN ICP (10 it, thr 0.5) vs 5M-point map"),
    "c1": dict(map_points=1_000_000, beams=32, azimuths=1024, leaf=0.0, mode="reference", max_range=100.0,
               desc="32-beam scan (~30k pts) reference point-to-point ICP (10 it, thr 0.5) vs 1M-point map"),
    "c3": dict(map_points=50_000_000, beams=32, azimuths=1024, leaf=0.0, mode="gn_p2plane", max_range=100.0,
               sharded=True, scans_per_step=16,
               desc="32-beam scans (~30k pts) point-to-plane GN ICP (10 it, thr 0.5) vs 50M-point map, "
                    "map sharded by cell columns across ranks, one 32-double all-reduce per scan per iteration"),
    # config 4: the offline sequence (10 000 scans of config 1's shape against the 5M-point map); scans are
    # independent, so ranks take disjoint scans and the sequence time is 10 000 / (scans/s over all ranks)
    "c4": dict(map_points=5_000_000, beams=32, azimuths=1024, leaf=0.0, mode="reference", max_range=100.0,
               scans_per_step=128,
               desc="offline reprocessing: 32-beam scans (~30k pts) reference point-to-point ICP (10 it, thr 0.5) vs "
                    "5M-point map, scans sharded across ranks, no communication"),
    # config 5 per GPU: 62.5M map points per rank (500M on 8), dense scans, tight leaf, 30 iterations.
    # (PCL's index-overflow guard refuses leaf 0.05 on a 100 m scan, SURVEY 7.3-6: range clamped to 40 m.)
    "c5": dict(map_points=62_500_000, per_rank=True, beams=128, azimuths=2048, leaf=0.05, mode="gn_p2p", max_range=40.0,
               sharded=True, scans_per_step=8, iters=30, normals=False,
               desc="128-beam dense scans (~260k rays) voxel 0.05 m + point-to-point GN ICP (30 it, thr 0.5) vs "
                    "62.5M map points per rank, map sharded by columns across ranks"),
    "mini": dict(map_points=200_000, beams=16, azimuths=512, leaf=0.2, mode="gn_p2plane", max_range=60.0,
                 desc="smoke-size variant of c2"),
}
THR, ITERS = 0.5, 10


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def make_workload(name: str, n_scans: int, rank: int, distinct: int = 16, world: int = 1):
    """Map + a batch of raw scans with perturbed initial poses (all seeded)."""
    from ssf_gpu import synth
    w = WORKLOADS[name]
    if w.get("sharded"):
        rank = 0  # map-sharded workloads: every rank registers the SAME scans against its map shard
    t0 = time.time()
    m_points = w["map_points"] * (world if w.get("per_rank") else 1)
    xyz, nrm, half = synth.make_map(m_points, normals=w.get("normals", True))
    log(f"[bench r{rank}] map {xyz.shape[0]} pts, half extent {half} m, {time.time() - t0:.1f}s")
    t0 = time.time()
    distinct = min(distinct, n_scans)
    base, scans, inits, gts = [], [], [], []
    for d in range(distinct):
        k = 1000 * rank + 40 * d + (int(half / 0.15) - 320 if w.get("sharded") else 0)  # sharded: map centre
        T = synth.street_pose(k, half=half)
        base.append((T, synth.make_scan(T, w["beams"], w["azimuths"], scan_id=k, max_range=w["max_range"])))
    for s in range(n_scans):
        T, sc = base[s % distinct]
        scans.append(sc)
        gts.append(T)
        inits.append(synth.perturb_pose(T, 100000 * rank + s))
    log(f"[bench r{rank}] {distinct} distinct scans of ~{scans[0].shape[0]} pts, {time.time() - t0:.1f}s")
    return w, xyz, nrm, half, scans, inits, gts


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).
    In-process NVML from a thread, every 2 ms -- the timed region of a default run lasts tens of
    milliseconds, too short for an `nvidia-smi -lms` child to report even once; nvidia-smi is the
    fallback when NVML cannot be loaded."""

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc, self.thread, self.stop_flag = gpu_index, [], None, None, False
        self.nvml, self.handle = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((float(sm), float(mx), int(rs)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        self.stop_flag = False
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._sample_nvml, daemon=True)
            self.thread.start()
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        self.stop_flag = True
        if self.nvml is not None:
            if self.thread:
                self.thread.join(timeout=1.0)
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [r[0] for r in self.rows]
            mx = max([r[1] for r in self.rows], default=0.0)
            reasons = sorted(k for k, b in bits.items() if any(r[2] & b for r in self.rows))
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 2 ms period, during the timed regions"}
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


def pose_delta(Ta, Tb):
    Ta, Tb = np.asarray(Ta, np.float64), np.asarray(Tb, np.float64)
    dt = float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))
    R = Ta[:3, :3].T @ Tb[:3, :3]
    return dt, float(np.arccos(max(-1.0, min(1.0, (np.trace(R) - 1.0) / 2.0))))


def nn_footprint_bytes(map_xyz, queries_world, cell):
    """SURVEY 8(d) algorithmic bytes of one search launch: every distinct map point (16 B) and
    cell entry (8 B) in the 3x3x3 neighbourhood of a query-occupied cell, counted once, plus
    20 B per query (16 B read, 4 B correspondence written)."""
    o = map_xyz[:, :3].min(0)
    # only map points near the queries can lie in a query cell's neighbourhood: drop the rest first
    qlo, qhi = queries_world.min(0) - 2 * cell, queries_world.max(0) + 2 * cell
    near = ((map_xyz[:, :3] >= qlo) & (map_xyz[:, :3] <= qhi)).all(1)
    map_xyz = map_xyz[near]
    mc = np.floor((map_xyz[:, :3] - o) / cell).astype(np.int64)
    dims = mc.max(0) + 3
    mkey = ((mc[:, 2] + 1) * dims[1] + (mc[:, 1] + 1)) * dims[0] + (mc[:, 0] + 1)
    ukeys, counts = np.unique(mkey, return_counts=True)
    qc = np.floor((queries_world - o) / cell).astype(np.int64)
    ok = ((qc >= -1) & (qc <= dims - 2)).all(1)
    qc = np.unique(qc[ok], axis=0)
    neigh = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                c = qc + np.array([dx, dy, dz])
                neigh.append(((c[:, 2] + 1) * dims[1] + (c[:, 1] + 1)) * dims[0] + (c[:, 0] + 1))
    nk = np.unique(np.concatenate(neigh))
    pos = np.searchsorted(ukeys, nk)
    pos[pos >= len(ukeys)] = len(ukeys) - 1
    hit = ukeys[pos] == nk
    n_pts, n_cells = int(counts[pos[hit]].sum()), int(hit.sum())
    return 20 * queries_world.shape[0] + 16 * n_pts + 8 * n_cells, n_pts, n_cells


# ------------------------------------------------------------------------------------------------
def run_cpu(args, rank, world):
    """--impl reference: the CPU restatement (oracle port) on a bounded sample, all host threads."""
    if rank != 0:
        return
    from oracle import oracle
    sample = max(1, min(args.cpu_scans, args.scans_per_step))
    w, xyz, nrm, half, scans, inits, gts = make_workload(args.workload, sample, 0, distinct=sample)
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: the oracle takes
    # its thread count per call, so the environment default does not cap it)
    threads = max(oracle.max_threads(), len(os.sched_getaffinity(0)))
    t0 = time.time()
    tree = oracle.KdTree(xyz)
    build_s = time.time() - t0

    def step():
        q = 0
        for sc, T0 in zip(scans, inits):
            src = oracle.voxel_grid(sc, w["leaf"])[0] if w["leaf"] > 0 else sc
            if w["mode"] in ("gn_p2plane", "gn_p2p"):
                r, _ = oracle.icp_gn(tree, src, T0, mode="p2plane" if w["mode"] == "gn_p2plane" else "p2p", normals=nrm,
                                     max_correspondence_dist=THR, num_iterations=w.get("iters", ITERS), threads=threads)
            else:
                r, _, _ = oracle.icp_reference(tree, src, T0, THR, ITERS, 0.05, 1e-5, threads=threads)
            q += src.shape[0] * r.n_searches
        return q

    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.time()
    queries = 0
    steps = max(1, min(args.steps, 3))
    for _ in range(steps):
        queries += step()
    dt = time.time() - t0
    val = sample * steps / dt
    line = {"impl": "reference", "metric": "icp_scans_per_sec", "value": val, "unit": "scans/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args, w),
            "nn_queries_per_sec": queries / dt,
            "cpu_baseline": {"value": val, "unit": "scans/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} scans/step x {steps} steps, KD-tree build {build_s:.1f}s excluded"},
            "e2e": {"value": val, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def config_of(args, w):
    return {"workload": f"{args.workload}: {w['desc']}", "scans_per_step": args.scans_per_step,
            "map_points": w["map_points"], "scan_rays": w["beams"] * w["azimuths"], "voxel_leaf": w["leaf"],
            "mode": w["mode"], "max_correspondence_dist": THR, "iterations": w.get("iters", ITERS),
            "parallelism": (f"map-sharded x{args.gpus} (scans replicated, per-iteration sum of 32 doubles per scan: " +
                            ("in-kernel exchange over peer memory" if args.exchange == "peer" else "NCCL all-reduce hook") + ")"
                            if w.get("sharded")
                            else f"scan-sharded x{args.gpus} (map replicated)"),
            "l2": args.l2_note}


def run_gpu(args, rank, world, local_rank):
    import torch
    import ssf_gpu
    from ssf_gpu import capi

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # NCCL prints its version banner on stdout; keep stdout for the one JSON line
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    B = args.scans_per_step
    w, xyz, nrm, half, scans, inits, gts = make_workload(args.workload, B, rank, world=world)
    iters = w.get("iters", ITERS)
    sharded = bool(w.get("sharded")) and world > 1
    ctx = ssf_gpu.Context(local_rank)
    mode = {"gn_p2plane": ssf_gpu.MODE_GN_P2PLANE, "gn_p2p": ssf_gpu.MODE_GN_P2P,
            "reference": ssf_gpu.MODE_REFERENCE}[w["mode"]]
    acc, eps = (0.05, 1e-5) if mode == ssf_gpu.MODE_REFERENCE else (0.0, 0.0)
    icp = ssf_gpu.ICPPointToPoint(THR, iters, acc, eps, mode=mode, reduce=ssf_gpu.REDUCE_STRICT, context=ctx)
    icp.setSourceVoxelLeaf(w["leaf"])
    t0 = time.time()
    if sharded:
        from ssf_gpu import shard
        sh = shard.shard_map(xyz, nrm, rank, world, THR)
        icp.setTargetShard(sh)
        if args.exchange == "peer":
            shard.setup_peer_exchange(icp, rank, world, max_scans=B)
        else:
            icp.setAllreduce(shard.torch_allreduce_hook(local_rank))
        log(f"[bench r{rank}] shard {sh['points'].shape[0]} of {xyz.shape[0]} pts, columns {sh['own']}")
    else:
        icp.setTargetPointCloud(xyz, nrm)
    log(f"[bench r{rank}] map index built in {time.time() - t0:.2f}s")

    n_pts = [s.shape[0] for s in scans]
    total = int(sum(n_pts))
    pinned = torch.empty((total, 4), dtype=torch.float32, pin_memory=True)
    pinned.numpy()[:] = np.concatenate(scans, axis=0)
    T_pinned = torch.empty((B, 16), dtype=torch.float32, pin_memory=True)
    T_pinned.numpy()[:] = np.stack([np.asarray(T, np.float32).T.reshape(16) for T in inits])
    res = (capi.IcpResult * B)()
    batch = ssf_gpu.Batch(icp, B, total + 1)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def e2e_step():
        batch.upload_ptr(pinned.data_ptr(), n_pts, 16)
        batch.set_initial_ptr(T_pinned.data_ptr())
        batch.run()
        batch.results_into(res)

    # ---- correctness gate on the measured configuration: results must be sane ----------------
    e2e_step()
    errs = [pose_delta(ssf_gpu._rowmajor(r.transformation), T)[0] for r, T in zip(res, gts)]
    n_src = [int(r.n_source) for r in res]
    searches = [int(r.n_searches) for r in res]
    log(f"[bench r{rank}] check: median |t - t_gt| = {np.median(errs):.4f} m, n_source ~{int(np.median(n_src))}, "
        f"searches {int(np.median(searches))}, iterations {int(np.median([r.iterations for r in res]))}")
    errs0 = [pose_delta(T0, T)[0] for T0, T in zip(inits, gts)]
    bar = 0.1 if w["mode"] != "reference" else float(np.median(errs0))  # the reference's lazy p2p loop converges slowly
    if not (np.median(errs) < bar):
        raise SystemExit("bench: registration did not converge on the benchmark workload")
    # single-scan latency through the per-object API (ssf_icp_align), device time by CUDA events
    icp.setSourcePointCloud(scans[0])
    icp.setInitialTransformation(inits[0])
    single_ms = float(np.median([icp.calculateAlignment().device_ms for _ in range(5)]))

    # ---- value: inputs resident in HBM ----------------------------------------------------------
    for _ in range(args.warmup):
        batch.run()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # inputs of one step: raw scans + map (+ normals).  Smaller than the 126 MB L2 -> flush L2 between
    # timed steps (write a 256 MB buffer); larger -> the step itself streams them
    input_bytes = total * 16 + xyz.shape[0] * 16 * (2 if w["mode"] == "gn_p2plane" else 1)
    if sharded:
        input_bytes = total * 16 + sh["points"].shape[0] * 16 * (2 if w["mode"] == "gn_p2plane" else 1)
    flush = input_bytes < 126e6
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}") if flush else None

    def timed_steps():
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for e0, e1 in evs:
            if flush:
                with torch.cuda.stream(stream):
                    flush_buf.zero_()
            e0.record(stream)
            batch.run()
            e1.record(stream)
        ctx.synchronize()
        evs[-1][1].synchronize()
        return float(sum(e0.elapsed_time(e1) for e0, e1 in evs))

    # timed region 1 -> `value`: K steps of the product path (ssf_batch_run as a user calls it: the
    # GN loop replays its CUDA graph where the launch sequence can be captured)
    launches0 = capi.lib().ssf_kernel_launches()
    dev_ms = timed_steps()
    launches = int(capi.lib().ssf_kernel_launches() - launches0)
    barrier()
    # timed region 2 -> `roofline`: the same K steps with every K3 launch bracketed by a pair of CUDA
    # events on the library's stream (plain launches -- a host-side event cannot sit inside the graph).
    # One untimed step first, so that the event pool exists before the region starts (creating 2 x K x
    # iterations events inside it cost the map-sharded workload a quarter of its step).
    ctx.time_searches(True)
    batch.run()
    ctx.search_time()
    barrier()
    dev_ms_timed = timed_steps()
    search_ms, search_launches = ctx.search_time()
    ctx.time_searches(False)
    if os.environ.get("SSF_BENCH_DEBUG"):
        log(f"[bench r{rank}] debug: {dev_ms / args.steps:.3f} ms/step; with search events {dev_ms_timed / args.steps:.3f} ms/step, "
            f"search {search_ms / max(1, search_launches) * 1e3:.1f} us/launch ({search_launches} launches)")
    batch.results_into(res)
    queries_per_step = int(sum(int(r.n_source) * int(r.n_searches) for r in res))
    barrier()

    # ---- e2e: host buffers in, host results out ---------------------------------------------------
    # Two batches in flight: the H2D copy of step i+1 (per-batch copy stream) overlaps the alignment
    # of step i, and the D2H of step i's results is collected while step i+1 runs.  Every step's
    # copies are inside the timed region.
    batch2 = ssf_gpu.Batch(icp, B, total + 1)
    pair = [batch, batch2]

    def e2e_pipelined(steps):
        pending = None
        for i in range(steps):
            b = pair[i % 2]
            b.upload_ptr(pinned.data_ptr(), n_pts, 16, wait=False)
            b.set_initial_ptr(T_pinned.data_ptr())
            b.run()
            if pending is not None:
                pending.results_into(res)
            pending = b
        pending.results_into(res)

    e2e_pipelined(max(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record(stream)
    e2e_pipelined(args.steps)
    ee1.record(stream)
    ctx.synchronize()
    ee1.synchronize()
    # the first H2D starts on the copy stream before ee0 completes on the compute stream: take the
    # larger of the device interval and the host wall clock around the same calls
    e2e_s = max(ee0.elapsed_time(ee1) * 1e-3, time.perf_counter() - t0)
    clocks = sampler.stop()
    barrier()

    t_dev = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = [float(x) for x in t_dev.cpu()]
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    job_scans = B if sharded else world * B  # map-sharded ranks cooperate on the same B scans
    value = job_scans * args.steps / (dev_ms_max * 1e-3)
    e2e_val = job_scans * args.steps / (e2e_ms_max * 1e-3)
    # roofline of the dominant kernel (search_accum): algorithmic bytes of ONE launch over the batch
    peak, peak_kind = peaks()
    cell = float(np.sqrt(np.float32(THR)) * np.float32(1.01))
    down = [ssf_gpu.voxel_down_sample(s, w["leaf"], ctx) if w["leaf"] > 0 else s[:, :3] for s in scans[:16]]
    qw = np.concatenate([d @ np.asarray(T, np.float64)[:3, :3].T + np.asarray(T, np.float64)[:3, 3]
                         for d, T in zip(down, inits[:16])])
    fp16, n_pts_fp, n_cells_fp = nn_footprint_bytes(xyz, qw, cell)
    # 16 of the B scans are distinct; the other scans revisit the same neighbourhoods
    q_per_launch = queries_per_step / max(1, int(np.median(searches)))
    alg_bytes = 20.0 * q_per_launch + 16.0 * n_pts_fp + 8.0 * n_cells_fp
    avg_search_ms = search_ms / max(1, search_launches)
    achieved = alg_bytes / (avg_search_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("workload") == args.workload and tj.get("scans_per_step") == B:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "search_accum_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_kind": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                "alg_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_search_ms,
                "share_of_step": search_ms / dev_ms_timed, "ms_per_step_with_events": dev_ms_timed / args.steps,
                "queries_per_launch": q_per_launch,
                "bytes_per_query": alg_bytes / max(1.0, q_per_launch)}
    args.l2_note = (f"per-step inputs {input_bytes / 1e6:.0f} MB " +
                    ("< 126 MB L2: L2 flushed (256 MB write) between timed steps" if flush
                     else "exceed the 126 MB L2: no flush needed"))
    line = {"metric": "icp_scans_per_sec", "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if w.get("sharded") else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(args, w),
            "nn_queries_per_sec": (1 if sharded else world) * queries_per_step * args.steps / (dev_ms_max * 1e-3),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "scans/s", "h2d_bytes_per_step": total * 16 + B * 64,
                    "d2h_bytes_per_step": B * ctypes.sizeof(capi.IcpResult)},
            "gpu_launches": launches, "single_scan_ms": single_ms, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, w, xyz, nrm, scans, inits)
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def cpu_baseline(args, w, xyz, nrm, scans, inits):
    """The oracle port on this box's host cores, bounded sample (rank 0, N = 1 only)."""
    from oracle import oracle
    threads = max(oracle.max_threads(), len(os.sched_getaffinity(0)))
    t0 = time.time()
    tree = oracle.KdTree(xyz)
    build_s = time.time() - t0
    n = max(1, min(args.cpu_scans, len(scans)))

    def run(th):
        t0 = time.time()
        for sc, T0 in zip(scans[:n], inits[:n]):
            src = oracle.voxel_grid(sc, w["leaf"])[0] if w["leaf"] > 0 else sc
            if w["mode"] in ("gn_p2plane", "gn_p2p"):
                oracle.icp_gn(tree, src, T0, mode="p2plane" if w["mode"] == "gn_p2plane" else "p2p", normals=nrm,
                              max_correspondence_dist=THR, num_iterations=w.get("iters", ITERS), threads=th)
            else:
                oracle.icp_reference(tree, src, T0, THR, ITERS, 0.05, 1e-5, threads=th)
        return n / (time.time() - t0)

    v_all = run(threads)
    v_one = run(1) if n <= 8 else None
    return {"value": v_all, "unit": "scans/s", "cores": threads, "kind": "port",
            "sample": f"{n} scans of the same workload, KD-tree build ({build_s:.1f}s) excluded",
            "value_single_thread": v_one}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scans-per-step", type=int, default=0, help="default: 64 (16 for c3)")
    ap.add_argument("--cpu-scans", type=int, default=8, help="scans in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="map-sharded workloads: in-kernel exchange over peer memory (default) or the NCCL all-reduce hook")
    args = ap.parse_args()
    args.l2_note = "n/a (CPU run)"
    if args.scans_per_step <= 0:
        args.scans_per_step = WORKLOADS[args.workload].get("scans_per_step", 64)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_cpu(args, rank, world)
        return
    if args.warmup < 3:
        log("[bench] note: timing rules ask for >= 3 warm-up steps")
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
