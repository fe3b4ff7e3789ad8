"""CPU: bench.py's reference arm prints the contract's JSON line, and both arms describe the workload with the same
``config`` object (the driver compares them)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == os.cpu_count()
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    import bench
    assert line["config"] == bench.WORKLOADS[line["config"]["workload"]]().config()


def test_workload_table():
    import bench
    assert set(bench.WORKLOADS) == {"canny_hough_1080p_b256", "canny_houghp_4k_b128", "lsd_1080p_b512",
                                    "lines_1080p_cave_b128"}
    w = bench.WORKLOADS["canny_hough_1080p_b256"]()
    assert w.chain_bytes() == 12664584
    assert bench.WORKLOADS["canny_houghp_4k_b128"]().chain_bytes() == 33177600
    assert bench.WORKLOADS["lsd_1080p_b512"]().chain_bytes() == 6220800
    for c in bench.WORKLOADS.values():
        json.dumps(c().config())


def test_seg_match_and_numa_binding_without_a_gpu():
    import numpy as np
    import bench
    a = np.array([[0, 0, 10, 10], [5, 5, 9, 1]], np.float32)
    assert bench.seg_match(a, a + 0.4, 0.5)
    assert bench.seg_match(a, a[:, [2, 3, 0, 1]] - 0.3, 0.5)      # either orientation
    assert not bench.seg_match(a, a + 0.6, 0.5)
    assert not bench.seg_match(a, a[:1], 0.5)                      # counts differ
    assert bench.seg_match(a[:0], a[:0], 0.5)
    import graphics_project_b200 as gp
    info = gp.pipeline.bind_to_gpu_numa_node(0)                    # no NVML / no GPU here: says so, changes nothing
    assert info["bound"] is False and ("error" in info or info["numa_node"] in (None, -1))


def test_reference_arm_other_workloads_share_the_engine_config():
    import bench
    for name in ("canny_houghp_4k_b128", "lsd_1080p_b512", "lines_1080p_cave_b128"):
        w = bench.WORKLOADS[name]()
        c = w.config()
        assert c["workload"] == name and c["batch_per_gpu"] == w.batch and c["height"] * c["width"] == w.h * w.w
        assert w.chain_bytes() > 0 and set(w.kernel_bytes())
