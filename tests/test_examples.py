"""Host-side callers next to the hot path: examples/pwpp_sequence.cpp, the demo_sequential
equivalent with a double-buffered page-locked reader. On the CPU its pipeline runs against a stub of the C-ABI
(tests/stub_pwpp.c); against the real library without a CUDA device it must fail loudly."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
SRC = os.path.join(REPO, "examples", "pwpp_sequence.cpp")


def _scans(tmp_path, kitti, sizes):
    for i, n in enumerate(sizes):
        np.ascontiguousarray(kitti[i % 6][:n]).tofile(tmp_path / f"{i:06d}.bin")
    (tmp_path / "notes.txt").write_text("not a scan")


def test_sequence_runner_pipeline_with_stub(tmp_path, kitti):
    build = os.path.join(HERE, "_build")
    os.makedirs(build, exist_ok=True)
    stub = os.path.join(build, "libpwpp_stub.so")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-I" + os.path.join(REPO, "include"), os.path.join(HERE, "stub_pwpp.c"), "-o", stub])
    exe = os.path.join(build, "pwpp_sequence_stub")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(REPO, "include"), SRC, "-o", exe, stub, "-Wl,-rpath," + build, "-lpthread"])
    sizes = [5000, 124000, 300, 60000, 7]     # growing and shrinking scans: the pinned buffers are re-sized on demand
    _scans(tmp_path, kitti, sizes)
    out = subprocess.run([exe, str(tmp_path), "--repeat", "3"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("0000")]
    assert [l.split()[0] for l in lines] == [f"{i:06d}.bin" for i in range(5)] * 3      # file-name order, every frame once per pass
    for l, n in zip(lines, sizes * 3):
        tok = l.split()
        assert int(tok[2]) == n and int(tok[4]) + int(tok[6]) == n
    expected_ground = [int((kitti[i % 6][:n, 2] < -1.5).sum()) for i, n in enumerate(sizes)]
    assert [int(l.split()[4]) for l in lines[:5]] == expected_ground
    assert "15 frames" in out.stdout.splitlines()[-1]
    # an empty directory is an error, not a silent success
    empty = tmp_path / "empty"; empty.mkdir()
    assert subprocess.run([exe, str(empty)], capture_output=True).returncode == 2


def test_sequence_runner_fails_loudly_without_a_gpu(tmp_path, kitti):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    exe = os.path.join(REPO, "patchwork-plusplus_b200", "lib", "pwpp_sequence")
    if not os.path.exists(exe):
        import build as pw_build
        pw_build.build_examples()
    _scans(tmp_path, kitti, [1000])
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "no CUDA device" in out.stderr and "no CPU path" in out.stderr


def test_pointcloud2_adaptor_with_stub(tmp_path, kitti):
    """include/patchwork/pointcloud2.hpp (the ROS 2 node's message handling without ROS) against the C-ABI
    stub: every message layout reaches the engine with the right coordinates (the stub labels by z), aligned equally
    spaced fields go through as a strided view without a gather, and the outgoing payloads are packed x/y/z."""
    build = os.path.join(HERE, "_build")
    os.makedirs(build, exist_ok=True)
    stub = os.path.join(build, "libpwpp_stub.so")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-I" + os.path.join(REPO, "include"), os.path.join(HERE, "stub_pwpp.c"), "-o", stub])
    exe = os.path.join(build, "pc2_driver_stub")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(REPO, "include"), os.path.join(HERE, "pc2_driver.cpp"), "-o", exe, stub,
                           "-Wl,-rpath," + build])
    a = kitti[2][:30000]
    np.ascontiguousarray(a).tofile(tmp_path / "scan.bin")
    out = subprocess.run([exe, str(tmp_path / "scan.bin")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    names = {"xyz12", "xyzi16", "pcl_xyzi32", "velodyne22", "ouster48_noint"}
    rows = {l.split()[0]: [int(x) for x in l.split()[1:]] for l in out.stdout.splitlines() if l.split() and l.split()[0] in names}
    ng = int((a[:, 2] < -1.5).sum())
    assert set(rows) == names
    for name, (zero_copy, g, n, payload, _chk) in rows.items():
        assert (g, n) == (ng, len(a) - ng), name
        assert payload == 12 * len(a), name
        assert zero_copy == (0 if name in ("pcl_xyzi32", "velodyne22") else 1), name
