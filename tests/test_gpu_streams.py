"""Stream-level features next to the hot path: state checkpoint / migration and zero-copy device results.
Runs after tests/test_gpu_parity.py (the parity gate) in the same `-m gpu` session."""
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
