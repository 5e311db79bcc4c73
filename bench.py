#!/usr/bin/env python
"""bench.py -- scan-to-map registration throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

Workload (BASELINE.json configs[1], "c2"): 64-beam x 2048-azimuth synthetic LiDAR scans
(~130k points) -> voxel-grid downsample, 0.2 m leaf -> point-to-plane Gauss-Newton ICP, 10
iterations, threshold 0.5, against a 5M-point synthetic map with analytic normals.
A step registers one batch of `--scans-per-step` DISTINCT scans (distinct poses); N ranks each take
their own batch against their own replica of the map (no collective on the data path: the
work shards by scan, "weak" scaling).

Printed keys: value = scans/s with the raw scans already resident in HBM (CUDA events on the
library's stream); e2e = the same through the host-buffer API (pinned host scans, packed 12-byte xyz
-> H2D -> align -> D2H results inside the timed region); roofline = the dominant kernel's algorithmic
bytes / its CUDA-event time against the measured HBM copy bandwidth; cpu_baseline = the CPU oracle on
a bounded sample of the same workload on this box's host cores; latency = single-scan device times;
map_sharded (N > 1) = the 50M-point map split across the N ranks with the per-iteration exchange
(BASELINE.json configs[2]), measured in the same run so the collective path is on record at every N.
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
    "c2": dict(map_points=5_000_000, beams=64, azimuths=2048, leaf=0.2, mode="gn_p2plane", max_range=100.0,
               scans_per_step=256,
               desc="64-beam scan (~130k pts) voxel 0.2 m + point-to-plane GN ICP (10 it, thr 0.5) vs 5M-point map"),
    "c1": dict(map_points=1_000_000, beams=32, azimuths=1024, leaf=0.0, mode="reference", max_range=100.0,
               scans_per_step=64,
               desc="32-beam scan (~30k pts) reference point-to-point ICP (10 it, thr 0.5) vs 1M-point map"),
    "c3": dict(map_points=50_000_000, beams=32, azimuths=1024, leaf=0.0, mode="gn_p2plane", max_range=100.0,
               sharded=True, scans_per_step=256,
               desc="32-beam scans (~30k pts) point-to-plane GN ICP (10 it, thr 0.5) vs 50M-point map, "
                    "map sharded by cell columns across ranks, one 32-double sum per scan per iteration"),
    # config 4: the offline sequence (10 000 scans of config 1's shape against the 5M-point map); scans are
    # independent, so ranks take disjoint scans and the sequence time is 10 000 / (scans/s over all ranks)
    "c4": dict(map_points=5_000_000, beams=32, azimuths=1024, leaf=0.0, mode="reference", max_range=100.0,
               scans_per_step=512,
               desc="offline reprocessing: 32-beam scans (~30k pts) reference point-to-point ICP (10 it, thr 0.5) vs "
                    "5M-point map, scans sharded across ranks, no communication"),
    # config 5 per GPU: 62.5M map points per rank (500M on 8), dense scans, tight leaf, 30 iterations.
    # (PCL's index-overflow guard refuses leaf 0.05 on a 100 m scan, SURVEY 7.3-6: range clamped to 40 m.)
    "c5": dict(map_points=62_500_000, per_rank=True, beams=128, azimuths=2048, leaf=0.05, mode="gn_p2p", max_range=40.0,
               sharded=True, scans_per_step=8, iters=30, normals=False,
               desc="128-beam dense scans (~260k rays) voxel 0.05 m + point-to-point GN ICP (30 it, thr 0.5) vs "
                    "62.5M map points per rank, map sharded by columns across ranks"),
    "mini": dict(map_points=200_000, beams=16, azimuths=512, leaf=0.2, mode="gn_p2plane", max_range=60.0,
                 scans_per_step=16, desc="smoke-size variant of c2"),
}
THR, ITERS = 0.5, 10
L2_BYTES = 126e6


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(name: str, n_scans: int, rank: int, world: int = 1, with_map: bool = True):
    """Map + a batch of n_scans DISTINCT raw scans (distinct poses) with perturbed initial poses, all seeded."""
    from ssf_gpu import synth
    w = WORKLOADS[name]
    seed_rank = 0 if w.get("sharded") else rank  # map-sharded: every rank registers the SAME scans against its shard
    t0 = time.time()
    m_points = w["map_points"] * (world if w.get("per_rank") else 1)
    if with_map and w.get("sharded") and world > 1 and m_points >= 100_000_000:
        # a map of hundreds of millions of points is synthesised ONCE (rank 0, all host threads) and handed to the
        # other ranks as a memory-mapped file; every rank then cuts its own shard out of it (shard.shard_map)
        path = os.path.join(os.environ.get("SSF_BENCH_TMP", "/tmp"), f"ssf_bench_map_{name}_{m_points}.npy")
        if rank == 0:
            synth.set_threads(len(os.sched_getaffinity(0)))  # the other ranks only wait meanwhile
            xyz, nrm, half = synth.make_map(m_points, normals=False)
            synth.set_threads(max(1, len(os.sched_getaffinity(0)) // world))
            np.save(path + ".tmp.npy", xyz)
            os.replace(path + ".tmp.npy", path)
            del xyz
        while not os.path.exists(path):
            time.sleep(0.5)
        xyz, nrm, half = np.load(path, mmap_mode="r"), None, synth.map_half_extent(m_points)[0]
    elif with_map:
        xyz, nrm, half = synth.make_map(m_points, normals=w.get("normals", True))
    else:
        xyz, nrm, half = None, None, synth.map_half_extent(m_points)[0]
    log(f"[bench r{rank}] {name}: map {m_points} pts, half extent {half} m, {time.time() - t0:.1f}s")
    t0 = time.time()
    scans, inits, gts = [], [], []
    per = int(4 * max(2.0, half - 15.0) / 0.15)  # poses of one lap of the route
    for s in range(n_scans):
        # poses spread over the whole route, a different stretch per rank
        k = (seed_rank * 7919 + s * max(1, per // max(1, n_scans)) + (per // 2 if w.get("sharded") else 0)) % per
        T = synth.street_pose(k, half=half)
        scans.append(synth.make_scan(T, w["beams"], w["azimuths"], scan_id=100000 * seed_rank + s, max_range=w["max_range"]))
        gts.append(T)
        inits.append(synth.perturb_pose(T, 100000 * seed_rank + s))
    log(f"[bench r{rank}] {name}: {n_scans} distinct scans of ~{scans[0].shape[0]} pts, {time.time() - t0:.1f}s")
    return w, xyz, nrm, half, scans, inits, gts


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).
    In-process NVML from a thread, every 2 ms; nvidia-smi is the fallback when NVML cannot be loaded."""

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
    """distinct map points (16 B) and cell entries (8 B) in the 3x3x3 neighbourhood (cell edge
    1.01 sqrt(thr)) of the query-occupied cells, each counted once.  Returns (points, cells)."""
    o = map_xyz[:, :3].min(0)
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
    return int(counts[pos[hit]].sum()), int(hit.sum())


def config_of(args, w):
    """Identical in both arms (ours / reference): the workload, nothing about how an arm sampled it."""
    return {"workload": f"{args.workload}: {w['desc']}", "scans_per_step": args.scans_per_step,
            "map_points": w["map_points"], "scan_rays": w["beams"] * w["azimuths"], "voxel_leaf": w["leaf"],
            "mode": w["mode"], "max_correspondence_dist": THR, "iterations": w.get("iters", ITERS),
            "parallelism": (f"map-sharded x{args.gpus} (scans replicated, per-iteration sum of 32 doubles per scan)"
                            if w.get("sharded") else f"scan-sharded x{args.gpus} (map replicated)")}


def host_threads():
    from oracle import oracle
    return max(oracle.max_threads(), len(os.sched_getaffinity(0)))


def oracle_scan(oracle, w, tree, nrm, sc, T0, threads):
    src = oracle.voxel_grid(sc, w["leaf"])[0] if w["leaf"] > 0 else sc
    if w["mode"] in ("gn_p2plane", "gn_p2p"):
        r, _ = oracle.icp_gn(tree, src, T0, mode="p2plane" if w["mode"] == "gn_p2plane" else "p2p", normals=nrm,
                             max_correspondence_dist=THR, num_iterations=w.get("iters", ITERS), threads=threads)
    else:
        r, _, _ = oracle.icp_reference(tree, src, T0, THR, ITERS, 0.05, 1e-5, threads=threads)
    return src.shape[0] * r.n_searches


# ------------------------------------------------------------------------------------------------
def run_cpu(args, rank, world):
    """--impl reference: the reference's algorithm on the host CPU, bounded sample, all host threads.
    Reference-loop workloads (c1, c4) run oracle/_ref -- the reference's own icp_point_to_point.cpp compiled
    unmodified (single-threaded, like the node) -- when it was shipped; the Gauss-Newton workloads have no
    counterpart in the reference and run the oracle port with OpenMP over the queries."""
    if rank != 0:
        return
    from oracle import oracle, ref
    w = WORKLOADS[args.workload]
    sample = max(1, min(args.cpu_scans, args.scans_per_step))
    w, xyz, nrm, half, scans, inits, gts = make_workload(args.workload, sample, 0)
    threads = host_threads()
    use_ref = w["mode"] == "reference" and ref.available()
    t0 = time.time()
    if use_ref:
        icp = ref.ICPPointToPoint(THR, ITERS, 0.05, 1e-5)
        icp.setDebugMode(False)
        icp.setTargetPointCloud(xyz)
        tree = None
    else:
        tree = oracle.KdTree(xyz)
    build_s = time.time() - t0

    def step():
        q = 0
        for sc, T0 in zip(scans, inits):
            if use_ref:
                icp.setSourcePointCloud(sc)
                icp.setInitialTransformation(T0)
                icp.calculateAlignment()
                q += sc.shape[0] * 5
            else:
                q += oracle_scan(oracle, w, tree, nrm, sc, T0, threads)
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
            "cpu_baseline": {"value": val, "unit": "scans/s", "cores": 1 if use_ref else threads,
                             "kind": "reference" if use_ref else "port",
                             "sample": f"{sample} scans/step x {steps} steps of the same workload, index build "
                                       f"({build_s:.1f}s) excluded"},
            "e2e": {"value": val, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(args, w, xyz, nrm, scans, inits):
    """The oracle port on this box's host cores, bounded sample (rank 0, N = 1 only), plus the same-lineage
    FLANN index (cv2.flann KDTREE_SINGLE) timed on one search pass of the sample's queries."""
    from oracle import oracle
    threads = host_threads()
    t0 = time.time()
    tree = oracle.KdTree(xyz)
    build_s = time.time() - t0
    n = max(1, min(args.cpu_scans, len(scans)))

    def run(th):
        t0 = time.time()
        for sc, T0 in zip(scans[:n], inits[:n]):
            oracle_scan(oracle, w, tree, nrm, sc, T0, th)
        return n / (time.time() - t0)

    v_all = run(threads)
    v_one = run(1) if n <= 8 else None
    out = {"value": v_all, "unit": "scans/s", "cores": threads, "kind": "port",
           "sample": f"{n} scans of the same workload, KD-tree build ({build_s:.1f}s) excluded",
           "value_single_thread": v_one}
    try:  # FLANN KDTREE_SINGLE (the index family PCL wraps), one core, one search pass
        import cv2
        sub = np.ascontiguousarray(xyz[:, :3])
        t0 = time.time()
        idx = cv2.flann_Index(sub, dict(algorithm=4, leaf_max_size=15))
        fb = time.time() - t0
        sc = scans[0]
        src = oracle.voxel_grid(sc, w["leaf"])[0] if w["leaf"] > 0 else sc
        T0 = np.asarray(inits[0], np.float32)
        q = np.ascontiguousarray((src[:, :3] @ T0[:3, :3].T + T0[:3, 3]).astype(np.float32))
        t0 = time.time()
        idx.knnSearch(q, 1, params=dict(checks=-1, eps=0.0, sorted=True))
        fq = time.time() - t0
        out["flann_kdtree_single"] = {"queries_per_sec": q.shape[0] / fq, "build_s": fb, "cores": 1,
                                      "queries": int(q.shape[0])}
    except Exception as e:  # cv2 absent on the box: say so instead of failing the bench
        out["flann_kdtree_single"] = {"unavailable": repr(e)[:120]}
    return out


def bind_to_gpu_numa(local_rank):
    """CPU affinity (and with it first-touch placement of the pinned staging buffers) on the NUMA node the
    rank's GPU hangs off.  Returns a description for the JSON line."""
    info = {"numa_node": None, "cpus": len(os.sched_getaffinity(0))}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        info["numa_node"] = node
        if node >= 0:
            cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
            ids = set()
            for part in cpus.split(","):
                a, _, b = part.partition("-")
                ids.update(range(int(a), int(b or a) + 1))
            ids &= os.sched_getaffinity(0)
            if ids:
                os.sched_setaffinity(0, ids)
                info["cpus"] = len(ids)
    except Exception as e:
        info["note"] = repr(e)[:100]
    return info


def run_gpu(args, rank, world, local_rank):
    import torch
    import ssf_gpu
    from ssf_gpu import capi

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        saved = os.dup(1)  # NCCL prints its version banner on stdout; keep stdout for the one JSON line
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
        elif args.exchange == "nccl":
            shard.setup_nccl(icp, rank, world)
        else:
            icp.setAllreduce(shard.torch_allreduce_hook(local_rank))
        log(f"[bench r{rank}] shard {sh['points'].shape[0]} of {xyz.shape[0]} pts, columns {sh['own']}")
        map_pts_dev = sh["points"].shape[0]
    else:
        icp.setTargetPointCloud(xyz, nrm)
        map_pts_dev = xyz.shape[0]
    log(f"[bench r{rank}] map index built in {time.time() - t0:.2f}s")

    n_pts = [s.shape[0] for s in scans]
    total = int(sum(n_pts))
    # host staging: PACKED 12-byte xyz (the ABI takes any stride >= 12; the 4th float of pcl::PointXYZ is padding)
    pinned = torch.empty((total, 3), dtype=torch.float32, pin_memory=True)
    pinned.numpy()[:] = np.concatenate([s[:, :3] for s in scans], axis=0)
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
        batch.upload_ptr(pinned.data_ptr(), n_pts, 12)
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
    input_bytes = total * 16 + map_pts_dev * 16 * (2 if w["mode"] == "gn_p2plane" else 1)
    flush = input_bytes < L2_BYTES
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}") if flush else None

    def timed_steps(n):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
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

    # timed region 1 -> `value`: K steps of the product path (ssf_batch_run as a user calls it)
    launches0 = capi.lib().ssf_kernel_launches()
    dev_ms = timed_steps(args.steps)
    launches = int(capi.lib().ssf_kernel_launches() - launches0)
    barrier()
    # timed region 2 -> `roofline`: steps with every K3 launch bracketed by a pair of CUDA events on the
    # library's stream (plain launches -- a host-side event cannot sit inside the graph).  One untimed
    # step first, so that the event pool exists before the region starts.
    k2 = max(1, min(args.steps, 20))
    ctx.time_searches(True)
    batch.run()
    ctx.search_time()
    barrier()
    dev_ms_timed = timed_steps(k2)
    per_launch = ctx.search_times(cap=8192)
    search_ms, search_launches = ctx.search_time()
    ctx.time_searches(False)
    batch.results_into(res)
    answered, walked = batch.search_stats() if w["mode"] != "reference" else (np.zeros(0), np.zeros(0))
    queries_per_step = int(answered.sum()) if answered.size else int(sum(int(r.n_source) * int(r.n_searches) for r in res))
    walked_per_step = int(walked.sum()) if walked.size else None
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
            b.upload_ptr(pinned.data_ptr(), n_pts, 12, wait=False)
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
    # what the host->device link gives this rank while every rank copies at once (pinned -> device, 5 x 256 MB)
    probe_src = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
    probe_dst = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    probe_dst.copy_(probe_src, non_blocking=True)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(5):
        probe_dst.copy_(probe_src, non_blocking=True)
    p1.record()
    p1.synchronize()
    h2d_gbs = 5 * (256 << 20) / (p0.elapsed_time(p1) * 1e-3) / 1e9
    del probe_src, probe_dst
    barrier()

    # ---- latency of the live-node case (config 1: one 32x1024 scan vs the 1M-point map) -----------------
    latency = {f"{args.workload}_single_scan_ms": single_ms}
    if rank == 0 and world == 1 and args.workload != "c1" and not args.no_latency:
        latency.update(c1_latency(ssf_gpu, ctx))

    # ---- the communicating multi-GPU path, on record at every N > 1 -------------------------------------
    map_sharded = None
    if world > 1 and not w.get("sharded") and not args.no_map_sharded:
        del batch2
        try:
            map_sharded = run_map_sharded(args, ssf_gpu, capi, torch, dist, ctx, rank, world, local_rank)
        except Exception as e:  # the headline line must still be printed
            map_sharded = {"error": repr(e)[:300]}

    t_dev = torch.tensor([dev_ms, e2e_s * 1e3, -h2d_gbs], dtype=torch.float64, device=f"cuda:{local_rank}")
    t_sum = torch.tensor([h2d_gbs], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_sum, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max, neg_min_h2d = [float(x) for x in t_dev.cpu()]
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    job_scans = B if sharded else world * B  # map-sharded ranks cooperate on the same B scans
    value = job_scans * args.steps / (dev_ms_max * 1e-3)
    e2e_val = job_scans * args.steps / (e2e_ms_max * 1e-3)
    peak, peak_kind = peaks()
    n_search = max(1, int(np.median(searches)))
    if w["mode"] == "reference":
        roofline = reference_roofline(peak, peak_kind, search_ms, search_launches, dev_ms_timed, k2, res, clocks, B, single_ms)
    else:
        # roofline of the dominant kernel (search_accum): algorithmic bytes of ONE launch over the batch
        cell = float(np.sqrt(np.float32(THR)) * np.float32(1.01))
        down = [ssf_gpu.voxel_down_sample(s, w["leaf"], ctx) if w["leaf"] > 0 else s[:, :3] for s in scans]
        qw = np.concatenate([d @ np.asarray(T, np.float64)[:3, :3].T + np.asarray(T, np.float64)[:3, 3]
                             for d, T in zip(down, inits)])
        n_pts_fp, n_cells_fp = nn_footprint_bytes(xyz, qw, cell)
        q_per_launch = queries_per_step / n_search
        alg_bytes = 20.0 * q_per_launch + 16.0 * n_pts_fp + 8.0 * n_cells_fp
        avg_search_ms = search_ms / max(1, search_launches)
        achieved = alg_bytes / (avg_search_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu --set full capture
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("workload") == args.workload and tj.get("scans_per_step") == B:
                traffic = tj.get("dram_bytes_per_launch")
        per_pos = [float(np.mean(per_launch[i::n_search])) for i in range(n_search)] if len(per_launch) >= n_search else []
        roofline = {"bound": "hbm", "kernel": "search_accum_kernel<GN_P2PLANE,128>" if w["mode"] == "gn_p2plane"
                    else "search_accum_kernel<GN_P2P,128>",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_kind": peak_kind, "alg_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_search_ms,
                    "launch_ms_by_iteration": per_pos,
                    "share_of_step": search_ms / dev_ms_timed, "ms_per_step_with_events": dev_ms_timed / k2,
                    "queries_per_launch": q_per_launch, "bytes_per_query": alg_bytes / max(1.0, q_per_launch),
                    "footprint_points": n_pts_fp, "footprint_cells": n_cells_fp}
    l2_note = (f"per-step inputs {input_bytes / 1e6:.0f} MB " +
               ("< 126 MB L2: L2 flushed (256 MB write) between timed steps" if flush
                else "exceed the 126 MB L2: no flush needed"))
    dev_s = dev_ms_max * 1e-3
    scale = 1 if sharded else world
    line = {"metric": "icp_scans_per_sec", "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if w.get("sharded") else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(args, w), "l2": l2_note,
            "nn_queries_per_sec": scale * queries_per_step * args.steps / dev_s,
            "nn_queries_walked_per_sec": (scale * walked_per_step * args.steps / dev_s) if walked_per_step is not None else None,
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "scans/s", "h2d_bytes_per_step": total * 12 + B * 64,
                    "d2h_bytes_per_step": B * ctypes.sizeof(capi.IcpResult), "point_stride_bytes": 12,
                    "h2d_probe_gbs_min_rank": -neg_min_h2d, "h2d_probe_gbs_sum": float(t_sum.cpu()[0]),
                    "rank0_numa": numa},
            "gpu_launches": launches, "single_scan_ms": single_ms, "latency": latency, "roofline": roofline}
    if map_sharded is not None:
        line["map_sharded"] = map_sharded
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, w, xyz, nrm, scans, inits)
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def reference_roofline(peak, peak_kind, search_ms, search_launches, dev_ms_timed, k2, res, clocks, B, single_ms):
    """REFERENCE-mode workloads (c1, c4): the dominant kernel is ref_reduce_kernel, a LATENCY-bound chain of
    dependent float adds in source-row order (what makes the result bit-identical to the reference), one
    block per scan; a batch runs the chains of all its scans side by side.  Its roofline is the 4.1-cycle
    FADD dependency floor per row and chain pass, not HBM bytes: `achieved` / `peak` are rows per cycle of ONE
    scan, taken from the single-scan device latency (so searches and launches are charged to it as well --
    an upper bound on the cycles per row).  The HBM fraction of the loop's bytes is reported beside it."""
    k_final = float(np.median([int(r.k_final) for r in res]))
    n_src = float(np.median([int(r.n_source) for r in res]))
    its = float(np.median([int(r.iterations) for r in res]))
    srch = float(np.median([int(r.n_searches) for r in res]))
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    chain_rows = n_src * (2.0 * its + (srch - 1.0))      # chain passes over the scan's rows: 2 per iteration + 1 per re-search
    cyc_per_row = single_ms * 1e-3 * mhz * 1e6 / max(1.0, chain_rows)
    loop_ms = (dev_ms_timed - search_ms) / k2            # per step: everything but the first search launch
    alg_bytes = B * its * 48.0 * k_final                  # 48 K bytes per loop iteration
    return {"bound": "latency (dependent FADD chain)", "kernel": "ref_reduce_kernel", "achieved": 1.0 / max(cyc_per_row, 1e-9),
            "peak": 1.0 / 4.1, "unit": "rows/cycle/scan", "frac": 4.1 / max(cyc_per_row, 1e-9), "traffic": None,
            "peak_kind": "4.1-cycle dependent FADD (profiles/exp/mb/chain.cu: 4.09 registers only, 4.19 fed from shared memory)",
            "cycles_per_row": cyc_per_row, "single_scan_ms": single_ms, "chain_passes_per_scan": 2.0 * its + (srch - 1.0),
            "hbm_frac_of_loop_bytes": alg_bytes / max(loop_ms * 1e-3, 1e-12) / 1e9 / peak, "hbm_peak_kind": peak_kind,
            "first_search_ms": search_ms / max(1, search_launches),
            "share_of_step": 1.0 - search_ms / dev_ms_timed, "ms_per_step_with_events": dev_ms_timed / k2}


def c1_latency(ssf_gpu, ctx):
    """Single-scan device latency of the live node's case (10 Hz budget, stochastic_filter.cpp:41): one 32x1024
    scan against the 1M-point map, in the reference's own loop (STRICT = bit-identical, FAST) and GN."""
    from ssf_gpu import synth
    xyz, nrm, half = synth.make_map(1_000_000, normals=True)
    T = synth.street_pose(100, half=half)
    scan = synth.make_scan(T, 32, 1024, scan_id=100)
    T0 = synth.perturb_pose(T, 100)
    out = {}
    icp = ssf_gpu.ICPPointToPoint(THR, ITERS, 0.05, 1e-5, context=ctx)
    icp.setTargetPointCloud(xyz, nrm)
    icp.setSourcePointCloud(scan)
    icp.setInitialTransformation(T0)
    for name, mode, red, acc, eps in (("c1_reference_strict_ms", ssf_gpu.MODE_REFERENCE, ssf_gpu.REDUCE_STRICT, 0.05, 1e-5),
                                      ("c1_reference_fast_ms", ssf_gpu.MODE_REFERENCE, ssf_gpu.REDUCE_FAST, 0.05, 1e-5),
                                      ("c1_gn_p2plane_ms", ssf_gpu.MODE_GN_P2PLANE, ssf_gpu.REDUCE_STRICT, 0.0, 0.0)):
        icp.setMode(mode, red)
        icp.setAcceptableMeanError(acc)
        icp.setTransformationEpsilon(eps)
        out[name] = float(np.median([icp.calculateAlignment().device_ms for _ in range(7)]))
    icp.close()
    return out


def run_map_sharded(args, ssf_gpu, capi, torch, dist, ctx, rank, world, local_rank):
    """BASELINE.json configs[2] in the same run: the 50M-point map split into `world` column shards (one-cell
    halo), the same scans on every rank, one sum of 32 doubles per scan per iteration -- through the in-kernel
    exchange over peer memory (`value`) and through ncclAllReduce behind the C ABI (`nccl`).  Strong scaling:
    rank 0 also registers the same scans against the WHOLE map on one GPU (`one_gpu_value`), the denominator."""
    from ssf_gpu import shard
    name = "c3"
    w = WORKLOADS[name]
    B = w["scans_per_step"]
    w, xyz, nrm, half, scans, inits, gts = make_workload(name, B, rank, world=world)
    icp = ssf_gpu.ICPPointToPoint(THR, ITERS, 0.0, 0.0, mode=ssf_gpu.MODE_GN_P2PLANE, context=ctx)
    t0 = time.time()
    sh = shard.shard_map(xyz, nrm, rank, world, THR)
    icp.setTargetShard(sh)
    build_s = time.time() - t0
    n_pts = [s.shape[0] for s in scans]
    total = int(sum(n_pts))
    cat = np.ascontiguousarray(np.concatenate([s[:, :3] for s in scans], axis=0))
    T0 = np.ascontiguousarray(np.stack([np.asarray(T, np.float32).T.reshape(16) for T in inits]))
    res = (capi.IcpResult * B)()
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    steps = max(5, min(args.steps, 30))

    def sync():
        ctx.synchronize()
        torch.cuda.synchronize()
        dist.barrier()

    def measure(batch):
        def timed(n):
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                batch.run()
            e1.record(stream)
            ctx.synchronize()
            e1.synchronize()
            return e0.elapsed_time(e1)

        for _ in range(3):
            batch.run()
        ms = timed(steps)
        batch.results_into(res)
        errs = [pose_delta(ssf_gpu._rowmajor(r.transformation), T)[0] for r, T in zip(res, gts)]
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.cpu()[0])
        # what an iteration spends outside the search kernel (row sum + exchange + solve + tile list): the
        # step with every search bracketed by events (plain launches), minus the searches
        ctx.time_searches(True)
        batch.run()
        ctx.search_time()
        ms_ev = timed(3)
        s_ms, s_n = ctx.search_time()
        ctx.time_searches(False)
        sync()
        return {"value": B * steps / (ms_max * 1e-3), "unit": "scans/s", "ms_per_step": ms_max / steps,
                "exchange_us_per_iter": 1e3 * (ms_ev - s_ms) / 3 / ITERS, "search_us_per_iter": 1e3 * s_ms / max(1, s_n),
                "median_pose_error_m": float(np.median(errs))}

    out = {"workload": f"c3: {w['desc']}", "n_ranks": world, "scans_per_step": B,
           "shard_points_rank0": int(sh["points"].shape[0]), "map_points": int(xyz.shape[0]), "shard_build_s_rank0": build_s}
    for ex in ("peer", "nccl"):
        batch = ssf_gpu.Batch(icp, B, total + 1)
        batch.upload_ptr(cat.ctypes.data, n_pts, 12)
        batch.set_initial_ptr(T0.ctypes.data)
        if ex == "peer":
            shard.setup_peer_exchange(icp, rank, world, max_scans=B)
        else:
            shard.setup_nccl(icp, rank, world)
        m = measure(batch)
        batch.close()
        if ex == "peer":
            icp.exchangeClose()
            out.update(m)
            out["exchange"] = "in-kernel over peer memory (rowsum_xchg_solve_kernel)"
        else:
            icp.ncclClose()
            out["nccl"] = m
    icp.close()
    if rank == 0:  # the denominator: the same scans against the whole map on one GPU
        one = ssf_gpu.ICPPointToPoint(THR, ITERS, 0.0, 0.0, mode=ssf_gpu.MODE_GN_P2PLANE, context=ctx)
        one.setTargetPointCloud(xyz, nrm)
        b1 = ssf_gpu.Batch(one, B, total + 1)
        b1.upload_ptr(cat.ctypes.data, n_pts, 12)
        b1.set_initial_ptr(T0.ctypes.data)
        for _ in range(3):
            b1.run()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            b1.run()
        e1.record(stream)
        ctx.synchronize()
        e1.synchronize()
        out["one_gpu_value"] = B * steps / (e0.elapsed_time(e1) * 1e-3)
        out["speedup_vs_one_gpu"] = out["value"] / out["one_gpu_value"]
        out["nccl"]["speedup_vs_one_gpu"] = out["nccl"]["value"] / out["one_gpu_value"]
        b1.close()
        one.close()
    dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scans-per-step", type=int, default=0, help="default: the workload's (256 for c2)")
    ap.add_argument("--cpu-scans", type=int, default=8, help="scans in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the config-1 single-scan latency block")
    ap.add_argument("--no-map-sharded", action="store_true", help="N > 1: skip the map-sharded (c3) block")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl", "hook"],
                    help="map-sharded workloads: in-kernel exchange over peer memory (default), ncclAllReduce behind the "
                         "C ABI, or the caller's all-reduce hook (torch.distributed)")
    args = ap.parse_args()
    if args.scans_per_step <= 0:
        args.scans_per_step = WORKLOADS[args.workload].get("scans_per_step", 64)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
        # torchrun's default for its workers; the scan / map generator and the oracle are OpenMP code and would
        # crawl on one thread: give every rank its share of the host cores (set before the libraries load)
        os.environ["OMP_NUM_THREADS"] = str(max(1, len(os.sched_getaffinity(0)) // int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    if args.impl == "reference":
        run_cpu(args, rank, world)
        return
    if args.warmup < 3:
        log("[bench] note: timing rules ask for >= 3 warm-up steps")
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
