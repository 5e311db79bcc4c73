"""Host-side pieces of bench.py (no GPU): the algorithmic-bytes footprint of a search launch, the workload table
and the committed DRAM-traffic record the roofline object quotes."""
import json
import os

import numpy as np

from conftest import ROOT

import bench


def _footprint_reference(map_xyz, q, cell):
    """SURVEY 8(d), literally: distinct map points and cells in the 3x3x3 neighbourhood of query-occupied cells."""
    o = map_xyz.min(0)
    mc = np.floor((map_xyz - o) / cell).astype(np.int64)
    qc = {tuple(c) for c in np.floor((q - o) / cell).astype(np.int64)}
    want = {(c[0] + dx, c[1] + dy, c[2] + dz) for c in qc for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)}
    cells = {}
    for c in map(tuple, mc):
        cells[c] = cells.get(c, 0) + 1
    hit = [c for c in cells if c in want]
    return 20 * len(q) + 16 * sum(cells[c] for c in hit) + 8 * len(hit)


def test_footprint_matches_the_definition():
    rng = np.random.default_rng(3)
    m = np.c_[rng.uniform(-30, 30, (20000, 2)), rng.normal(0, 0.05, 20000)].astype(np.float32)
    q = np.c_[rng.uniform(-5, 8, (500, 2)), rng.normal(0.2, 0.3, 500)].astype(np.float32)
    cell = 0.7142
    n_pts, n_cells = bench.nn_footprint_bytes(np.c_[m, np.ones(len(m), np.float32)], q, cell)
    assert 20 * len(q) + 16 * n_pts + 8 * n_cells == _footprint_reference(m, q, cell)
    assert 0 < n_pts < len(m) and n_cells > 0


def test_workloads_name_the_baseline_configs():
    cfg = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert len(cfg) == 5
    w = bench.WORKLOADS
    assert w["c1"]["mode"] == "reference" and w["c1"]["map_points"] == 1_000_000          # configs[0]
    assert w["c2"]["leaf"] == 0.2 and w["c2"]["mode"] == "gn_p2plane" and w["c2"]["map_points"] == 5_000_000  # configs[1]
    assert w["c3"]["sharded"] and w["c3"]["map_points"] == 50_000_000                     # configs[2]
    assert w["c4"]["mode"] == "reference" and w["c4"]["map_points"] == 5_000_000 and not w["c4"].get("sharded")  # configs[3]
    assert w["c5"]["leaf"] == 0.05 and w["c5"]["iters"] == 30 and w["c5"]["sharded"]      # configs[4], per-GPU scale


def test_traffic_record_matches_the_bench_workload():
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert t["workload"] == "c2" and t["scans_per_step"] == bench.WORKLOADS["c2"]["scans_per_step"]
    assert 1e6 < t["dram_bytes_per_launch"] < 1e9 and os.path.exists(os.path.join(ROOT, t["source"].split(":")[0]))


def test_both_arms_print_the_same_config():
    """The driver compares `config` of the two arms: nothing arm-specific (sample sizes, L2 notes) may sit in it."""
    import argparse
    a = argparse.Namespace(workload="c2", scans_per_step=256, gpus=1, exchange="peer")
    c = bench.config_of(a, bench.WORKLOADS["c2"])
    assert set(c) == {"workload", "scans_per_step", "map_points", "scan_rays", "voxel_leaf", "mode",
                      "max_correspondence_dist", "iterations", "parallelism"}
    assert c == bench.config_of(a, bench.WORKLOADS["c2"])


def test_distinct_scans_per_step():
    w, xyz, nrm, half, scans, inits, gts = bench.make_workload("mini", 12, 0)
    assert len(scans) == 12
    poses = {tuple(np.round(T[:3, 3], 3)) for T in gts}
    assert len(poses) == 12          # every scan of a step comes from its own pose
    w2, *_rest, gts2 = bench.make_workload("mini", 12, 1)
    assert {tuple(np.round(T[:3, 3], 3)) for T in gts2} != poses   # and every rank from its own stretch
