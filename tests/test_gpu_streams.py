"""Stream-level features next to the hot path: state checkpoint / migration and zero-copy device results.
Runs after tests/test_gpu_parity.py (the parity gate) in the same `-m gpu` session."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(num_streams=1):
    import pwpp_b200
    return pwpp_b200.Engine(device=0, num_streams=num_streams)


def test_state_export_import_migrates_a_stream(kitti):
    """Frames 0..2 on ctx A; the stream's blob is imported into stream 1 of a fresh ctx B; frame 3 on both gives
    bit-identical index lists, patch records, adaptive state and histories (a migrated sensor stream continues exactly)."""
    a = _engine()
    for f in range(3):
        a.estimate_host([kitti[f]])
    blob = a.export_state(0)
    b = _engine(num_streams=2)
    b.import_state(1, blob)
    a.estimate_host([kitti[3]])
    b.estimate_host([kitti[5], kitti[3]])    # stream 0 of B is an unrelated fresh stream
    assert np.array_equal(a.ground_indices(0), b.ground_indices(1)) and np.array_equal(a.nonground_indices(0), b.nonground_indices(1))
    assert bytes(a.bin_results(0)) == bytes(b.bin_results(1))
    sa, sb = a.state(0), b.state(1)
    assert bytes(sa) == bytes(sb)
    for r in range(4):
        for w in (0, 1):
            assert np.array_equal(a.history(0, r, w), b.history(1, r, w))
    # and the fresh neighbour stream was not disturbed
    c = _engine(); c.estimate_host([kitti[5]])
    assert np.array_equal(c.ground_indices(0), b.ground_indices(0))
    # the exported state differs from a fresh one (the test would be vacuous otherwise) and a wrong-size blob is refused
    assert blob != _engine().export_state(0)
    import pwpp_b200
    with pytest.raises(pwpp_b200.PwppError):
        b.import_state(0, blob[:-8])
    with pytest.raises(pwpp_b200.PwppError):
        b.import_state(0, b"\0" * len(blob))


def test_device_index_lists_are_the_host_getters(kitti):
    """pwpp_device_results through __cuda_array_interface__: the device-side lists equal what the copy_* getters return."""
    import torch
    frames = [kitti[0], kitti[1][:30000], np.zeros((0, 4), np.float32), kitti[2]]
    eng = _engine(num_streams=len(frames))
    eng.estimate_host(frames)
    eng.synchronize()
    idx, ng = eng.device_index_lists()
    assert idx.is_cuda and idx.dtype == torch.int32 and ng.dtype == torch.int32
    idx, ng = idx.cpu().numpy(), ng.cpu().numpy()
    offs = np.cumsum([0] + [len(f) for f in frames])
    for f in range(len(frames)):
        g, n = eng.ground_indices(f), eng.nonground_indices(f)
        assert ng[f] == len(g)
        assert np.array_equal(idx[offs[f]:offs[f] + len(g)], g)
        assert np.array_equal(idx[offs[f] + len(g):offs[f] + len(g) + len(n)], n)


def test_sequence_runner_matches_the_engine(tmp_path, kitti):
    """examples/pwpp_sequence.cpp (double-buffered page-locked reader -> PatchWorkpp drop-in class) over three scans twice:
    per-frame counts and adaptive sensor height equal a stream fed the same frames through the C-ABI."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(os.path.dirname(here), "patchwork-plusplus_b200", "lib", "pwpp_sequence")
    if not os.path.exists(exe):
        import build as pw_build
        pw_build.build_examples()
    for f in range(3):
        np.ascontiguousarray(kitti[f]).tofile(tmp_path / f"{f:06d}.bin")
    out = subprocess.run([exe, str(tmp_path), "--repeat", "2"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.splitlines() if l.startswith("0000")]
    assert len(lines) == 6
    eng = _engine()
    for t, tok in enumerate(lines):
        eng.estimate_host([kitti[t % 3]])
        assert int(tok[2]) == len(kitti[t % 3])
        assert int(tok[4]) == eng.num_ground(0) and int(tok[6]) == eng.num_nonground(0) and int(tok[8]) == eng.num_patches(0)
        assert abs(float(tok[10]) - eng.height(0)) < 1e-4


def test_narrow_ring_geometry_takes_the_exact_binning_kernel(kitti):
    """A concentric-zone layout with rings narrower than 1.5 m (zone 0: 5 rings of 0.875 m) is binned by the exact
    double-precision kernel (k_bin_hist<false>) instead of the fp32 filter: bin ids bit-exact and index sets identical
    to the oracle, including points a hair outside min_range / max_range."""
    import oracle_py as O
    from pwpp_ctypes import default_params
    import pwpp_b200
    p = default_params()
    p.min_range, p.max_range = 5.0, 40.0
    p.num_rings_each_zone[:] = [5, 1, 2, 3]
    p.num_sectors_each_zone[:] = [32, 54, 32, 32]
    extra = []
    for r in (5.0, 40.0):
        for dr in (0.0, 1e-5, -1e-5, 1.9e-4, -1.9e-4, 3e-4, -3e-4):
            for k in range(0, 54, 2):
                th = k * (2 * np.pi / 54) + 0.013
                extra.append([(r + dr) * np.cos(th), (r + dr) * np.sin(th), -1.7, 0.5])
    a = np.concatenate([kitti[0], np.array(extra, np.float32)])
    eng = pwpp_b200.Engine(p, device=0)
    eng.estimate_host([a])
    orc = O.Oracle(p, O.ARITH_CANON64); orc.estimate(a)
    assert np.array_equal(orc.bin_ids(), eng.bin_ids(0))
    if not (orc.bin_min_fit_n() < 3).any():
        assert np.array_equal(np.sort(orc.getGroundIndices()), np.sort(eng.ground_indices(0)))
        assert np.array_equal(np.sort(orc.getNongroundIndices()), np.sort(eng.nonground_indices(0)))


def test_pointcloud2_front_end_drives_the_real_engine(kitti, tmp_path):
    """include/patchwork/pointcloud2.hpp (the ROS 2 node's message handling, reference ros/src/GroundSegmentationServer.cpp:74-95,
    ros/src/Utils.hpp:158-195) against the REAL engine: PointCloud2-shaped buffers with point_step 12 / 16 / 32 / 22 / 48, with
    and without an intensity field, under the ROS launch-file parameters (ros/launch/patchworkpp.launch.py:50-64): every layout
    gives the same ground set as the engine fed with the packed N x 3 points."""
    import subprocess
    import pwpp_b200
    from param_sets import PARAM_SETS
    exe = os.path.join(os.path.dirname(pwpp_b200.LIB_PATH), "pc2_driver")
    assert os.path.exists(exe), "lib/pc2_driver was not built (patchwork-plusplus_b200/build.py)"
    a = np.ascontiguousarray(kitti[2][:60000])
    a.tofile(tmp_path / "scan.bin")
    out = subprocess.run([exe, str(tmp_path / "scan.bin"), "ros"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    names = {"xyz12", "xyzi16", "pcl_xyzi32", "velodyne22", "ouster48_noint"}
    rows = {l.split()[0]: [int(x) for x in l.split()[1:]] for l in out.stdout.splitlines() if l.split() and l.split()[0] in names}
    assert set(rows) == names
    mk, _ = PARAM_SETS["ros"]
    eng = pwpp_b200.Engine(mk(), device=0)
    eng.estimate_host([np.ascontiguousarray(a[:, :3])])
    g = eng.ground_indices(0).astype(np.int64)
    chk = int((g * (g % 97 + 1)).sum())
    for name, (zero_copy, ng, nn, payload, c) in rows.items():
        assert (ng, nn) == (len(g), len(a) - len(g)), name     # RNR is off in the launch file: intensity changes nothing
        assert c == chk, f"{name}: ground set differs from the engine's"
        assert payload == 12 * len(a), name
        assert zero_copy == (0 if name in ("pcl_xyzi32", "velodyne22") else 1), name


def test_pybind_device_tensors(kitti):
    """pypatchworkpp.estimateGround on CUDA tensors (__cuda_array_interface__ and DLPack, N x 4 and N x 3) and the device index
    views: identical to the numpy path, no host copy of the cloud."""
    import torch
    import pypatchworkpp as m
    a = kitti[1]
    P = m.Parameters(); P.verbose = False
    ref = m.patchworkpp(P); ref.estimateGround(a)
    g_ref, n_ref = ref.getGroundIndices(), ref.getNongroundIndices()
    t4 = torch.from_numpy(a).cuda()
    pw = m.patchworkpp(P)
    pw.estimateGround(t4)                                   # __cuda_array_interface__
    assert np.array_equal(pw.getGroundIndices(), g_ref) and np.array_equal(pw.getNongroundIndices(), n_ref)
    gd = torch.as_tensor(pw.getGroundIndicesDevice(), device="cuda")
    nd = torch.as_tensor(pw.getNongroundIndicesDevice(), device="cuda")
    assert gd.dtype == torch.int32 and np.array_equal(gd.cpu().numpy(), g_ref) and np.array_equal(nd.cpu().numpy(), n_ref)

    class OnlyDLPack:   # an object that offers DLPack but no __cuda_array_interface__
        def __init__(self, t): self.t = t
        def __dlpack__(self, stream=None): return self.t.__dlpack__()
        def __dlpack_device__(self): return self.t.__dlpack_device__()
    pw2 = m.patchworkpp(P)
    pw2.estimateGround(OnlyDLPack(t4))
    assert np.array_equal(pw2.getGroundIndices(), g_ref)
    # N x 3 on the device (the ROS node's input) vs N x 3 on the host
    P3 = m.Parameters(); P3.verbose = False; P3.enable_RNR = False
    h3, d3 = m.patchworkpp(P3), m.patchworkpp(P3)
    h3.estimateGround(np.ascontiguousarray(a[:, :3]))
    d3.estimateGround(torch.from_numpy(np.ascontiguousarray(a[:, :3])).cuda())
    assert np.array_equal(h3.getGroundIndices(), d3.getGroundIndices())
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        t = torch.from_numpy(a).cuda()
        pw4 = m.patchworkpp(P)                               # fresh temporal state, like `ref`
        pw4.estimateGround(t, stream=s.cuda_stream)         # enqueued on the producer's stream
        gd2 = torch.as_tensor(pw4.getGroundIndicesDevice(), device="cuda").clone()
    s.synchronize()
    assert np.array_equal(gd2.cpu().numpy(), g_ref)


def test_small_call_path_matches_the_batch_path(kitti, monkeypatch):
    """Calls of at most PWPP_SMALL_CALL frames take other kernels than batches (one CTA per patch above 512 points, the stand-alone
    front-end kernels): the same two scans through both paths give identical bin ids and index lists, patch planes within the
    GPU-vs-oracle tolerance, the same adaptive state; pwpp_call_times_us reports the device-side split of the one-chunk call."""
    frames = [kitti[0], kitti[4]]
    a = _engine(num_streams=2)                     # default: 2 frames <= 4 -> small-call kernels
    a.estimate_host(frames)
    ct = a.call_times_us()
    assert ct["device_total"] > 0 and abs(ct["h2d"] + ct["kernels"] + ct["d2h"] - ct["device_total"]) < 0.05 * ct["device_total"] + 5.0
    monkeypatch.setenv("PWPP_SMALL_CALL", "0")     # read when a context is created
    b = _engine(num_streams=2)
    b.estimate_host(frames)
    rec = np.dtype([("d", np.float64, 10), ("n", np.int32), ("ng", np.int32), ("verdict", np.int32), ("fitted", np.int32)])
    for f in range(2):
        assert np.array_equal(a.bin_ids(f), b.bin_ids(f))
        assert np.array_equal(a.ground_indices(f), b.ground_indices(f)) and np.array_equal(a.nonground_indices(f), b.nonground_indices(f))
        ra, rb = np.frombuffer(a.bin_results(f), rec), np.frombuffer(b.bin_results(f), rec)
        assert np.array_equal(ra["n"], rb["n"]) and np.array_equal(ra["ng"], rb["ng"]) and np.array_equal(ra["verdict"], rb["verdict"])
        fit = (ra["fitted"] != 0) & (ra["n"] >= 3)
        assert np.abs(ra["d"][fit][:, :6] - rb["d"][fit][:, :6]).max() <= 2e-9      # mean, normal
        assert abs(a.height(f) - b.height(f)) <= 1e-9
