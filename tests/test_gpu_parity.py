"""Parity of the CUDA path (through the C-ABI, lib/libpwpp_b200.so) with the oracle (CANON64 arithmetic).

Bar: polar bin ids bit-exact; ground / non-ground index SETS identical; per-patch plane parameters
and adaptive state within the stated double-precision tolerances (helpers.TOL_*). Patches in which the algorithm
itself is numerically undefined (a plane fitted to < 3 points: rank-deficient covariance, see
test_oracle_golden.py) are excluded from label comparison — with the default parameters that is ~0.1 % of patches.
"""
import numpy as np
import pytest

import oracle_py as O
from helpers import TOL_MEAN, TOL_NORMAL, assert_bins_close, assert_sets_equal, assert_state_close
from param_sets import PARAM_SETS

pytestmark = pytest.mark.gpu


def _engine(params=None, num_streams=1):
    import pwpp_b200
    return pwpp_b200.Engine(params, device=0, num_streams=num_streams)


def compare_frame(eng, f, orc, a, what, min_pts_floor=0):
    """Engine stream/frame f vs the oracle that has just processed the same points `a`."""
    n = a.shape[0]
    nb = orc.nbins
    ids_o, ids_e = orc.bin_ids(), eng.bin_ids(f)
    assert np.array_equal(ids_o, ids_e), f"{what}: {int((ids_o != ids_e).sum())} polar bin ids differ"
    bo, be = orc.bin_results(), eng.bin_results(f)
    degenerate = orc.bin_min_fit_n() < 3
    if min_pts_floor:
        degenerate |= np.array([bo[b].n < min_pts_floor for b in range(nb)])
    g_e, ng_e = eng.ground_indices(f), eng.nonground_indices(f)
    g_o, ng_o = orc.getGroundIndices(), orc.getNongroundIndices()
    assert len(g_e) + len(ng_e) == len(g_o) + len(ng_o), f"{what}: total emitted {len(g_e) + len(ng_e)} vs {len(g_o) + len(ng_o)}"
    allidx = np.concatenate([g_e, ng_e])
    assert len(np.unique(allidx)) == len(allidx), f"{what}: an index was emitted twice"
    if not degenerate.any():
        assert_sets_equal(g_o, ng_o, g_e, ng_e, n, what)
        assert_bins_close(bo, be, nb, what)
        assert eng.num_patches(f) == orc._f("num_patches")(orc._h)
        assert np.abs(eng.centers(f).astype(np.float64) - orc.getCenters()).max(initial=0) <= 1e-6, what
        assert np.abs(eng.normals(f).astype(np.float64) - orc.getNormals()).max(initial=0) <= 1e-6, what
        return 0
    # label comparison outside the degenerate patches
    pad = np.r_[degenerate, np.zeros(3, bool)]
    keep = ~pad[ids_o]
    mo = np.zeros(n, bool); mo[g_o] = True
    me = np.zeros(n, bool); me[g_e] = True
    assert np.array_equal(mo[keep], me[keep]), f"{what}: {int((mo[keep] != me[keep]).sum())} labels differ outside degenerate patches"
    for b in range(nb):
        if degenerate[b] or not bo[b].fitted or bo[b].n == 0:
            continue
        assert (bo[b].n, bo[b].n_ground) == (be[b].n, be[b].n_ground), f"{what} bin {b}"
        for k in range(3):
            assert abs(bo[b].normal[k] - be[b].normal[k]) <= TOL_NORMAL and abs(bo[b].mean[k] - be[b].mean[k]) <= TOL_MEAN, f"{what} bin {b}"
    return int(degenerate.sum())


@pytest.mark.parametrize("pname", list(PARAM_SETS))
def test_fixtures_fresh_batched(kitti, pname):
    """All six fixture scans in ONE call, each on its own fresh stream."""
    mk, cols = PARAM_SETS[pname]
    eng = _engine(mk(), num_streams=6)
    frames = [a[:, :cols] for a in kitti]
    eng.estimate_host(frames)
    for f, a in enumerate(frames):
        orc = O.Oracle(mk(), O.ARITH_CANON64)
        orc.estimate(a)
        nd = compare_frame(eng, f, orc, a, f"{pname}/fresh/{f}", min_pts_floor=5 if pname == "ros" else 0)
        if not nd:
            assert_state_close(eng.state(f), orc.state(), f"{pname}/fresh/{f}")


@pytest.mark.parametrize("pname", ["default", "no_rvpf_tgr"])
def test_fixtures_sequence(kitti, pname):
    """One stream over the six scans in order: exercises the temporal state (thresholds, histories, sensor height)."""
    mk, cols = PARAM_SETS[pname]
    eng = _engine(mk())
    orc = O.Oracle(mk(), O.ARITH_CANON64)
    diverged = False
    for f, a in enumerate(kitti):
        a = a[:, :cols]
        eng.estimate_host([a]); orc.estimate(a)
        nd = compare_frame(eng, 0, orc, a, f"{pname}/seq/{f}")
        diverged |= nd > 0 and pname != "default"
        if not diverged:
            assert_state_close(eng.state(0), orc.state(), f"{pname}/seq/{f}")
            for r in range(4):
                for w in (0, 1):
                    assert np.allclose(eng.history(0, r, w), orc.history(r, w), rtol=1e-6, atol=1e-9)
    if pname == "default":
        assert not diverged


def test_golden_sets_default(kitti, golden):
    """Against the committed golden vectors produced by the reference's own code (no oracle involved)."""
    eng = _engine(num_streams=6)
    eng.estimate_host(kitti)
    for f, a in enumerate(kitti):
        m = np.zeros(a.shape[0], bool); m[eng.ground_indices(f)] = True
        gm = np.unpackbits(golden[f"default/fresh/{f}/ground_mask"])[:a.shape[0]].astype(bool)
        assert np.array_equal(m, gm), f"frame {f}: {int((m != gm).sum())} labels differ from the reference golden"
        assert np.abs(eng.normals(f).astype(np.float64) - golden[f"default/fresh/{f}/normals"]).max() <= 1e-4
        assert np.abs(eng.centers(f).astype(np.float64) - golden[f"default/fresh/{f}/centers"]).max() <= 5e-5
        gs = golden[f"default/fresh/{f}/state"]
        st = eng.state(f)
        assert abs(st.sensor_height - gs[0]) <= 1e-5 and np.abs(np.array(st.elevation_thr) - gs[1:5]).max() <= 1e-5
    eng1 = _engine()
    for f, a in enumerate(kitti):
        eng1.estimate_host([a])
        m = np.zeros(a.shape[0], bool); m[eng1.ground_indices(0)] = True
        gm = np.unpackbits(golden[f"default/seq/{f}/ground_mask"])[:a.shape[0]].astype(bool)
        assert np.array_equal(m, gm), f"seq frame {f}: {int((m != gm).sum())} labels differ from the reference golden"


def test_synthetic_batch_vs_oracle():
    import synth
    nf = 24
    frames = [synth.make_frame(20260922, f).numpy() for f in range(nf)]
    eng = _engine(num_streams=nf)
    eng.estimate_host(frames)
    ndeg = 0
    for f, a in enumerate(frames):
        orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(a)
        ndeg += compare_frame(eng, f, orc, a, f"synthetic/{f}")
    assert ndeg <= 3


def test_streaming_synthetic_sequences():
    """4 streams x 5 consecutive frames each, processed as 5 batched calls."""
    import synth
    S, T = 4, 5
    eng = _engine(num_streams=S)
    orcs = [O.Oracle(arith=O.ARITH_CANON64) for _ in range(S)]
    for t in range(T):
        frames = [synth.make_frame(777 + s, t).numpy() for s in range(S)]
        eng.estimate_host(frames)
        for s in range(S):
            orcs[s].estimate(frames[s])
            nd = compare_frame(eng, s, orcs[s], frames[s], f"stream{s}/t{t}")
            assert nd == 0
            assert_state_close(eng.state(s), orcs[s].state(), f"stream{s}/t{t}")


def test_edge_cases():
    rng = np.random.default_rng(11)
    big_bin = np.c_[5 + rng.random(20000) * 0.5, rng.random(20000) * 0.5, -1.7 + rng.normal(0, 0.02, 20000), rng.random(20000)].astype(np.float32)
    nonfinite = np.array([[5, 1, np.nan, .5], [np.nan, 1, -1.7, .5], [5, np.inf, -1.7, .5], [6, 1, -np.inf, .5], [6, 1, np.inf, .01],
                          [7, 2, -1.7, np.nan]] + [[5 + 0.01 * i, 1 + rng.random() * 0.3, -1.7 + rng.normal(0, 0.01), .5] for i in range(30)], np.float32)
    # (filler points are scattered, not collinear: a rank-1 covariance has no defined normal, in the reference either)
    tomb = np.array([[5, 1, np.finfo(np.float32).tiny, .5]] + [[5 + 0.01 * i, 1.2 + rng.random() * 0.3, -1.7 + rng.normal(0, 0.01), .5] for i in range(15)], np.float32)
    rnr = np.array([[4, 0, -3.0, 0.05], [4, 0.1, -3.0, 0.5], [4, 0.2, -2.4, 0.05], [40, 0.2, -3.0, 0.05]] +
                   [[5 + 0.01 * i, 1 + rng.random() * 0.3, -1.7 + rng.normal(0, 0.01), .5] for i in range(12)], np.float32)
    cases = {
        "empty": np.zeros((0, 4), np.float32),
        "one_point": np.array([[5, 0, -1.7, 0.5]], np.float32),
        "nine_in_one_bin": np.c_[5 + rng.random(9) * 0.1, rng.random(9) * 0.1, -1.7 + rng.random(9) * 0.01, rng.random(9)].astype(np.float32),
        "ten_in_one_bin": np.c_[5 + rng.random(10) * 0.1, rng.random(10) * 0.1, -1.7 + rng.random(10) * 0.01, rng.random(10)].astype(np.float32),
        "all_out_of_range": np.c_[rng.random((50, 2)) * 1.0, rng.random((50, 2))].astype(np.float32),
        "flat_plane": np.c_[(rng.random((5000, 2)) - 0.5) * 60, np.full(5000, -1.723), rng.random(5000)].astype(np.float32),
        "axis_points": np.array([[10, 0, -1.7, .5], [-10, 0, -1.7, .5], [0, 10, -1.7, .5], [0, -10, -1.7, .5], [3, -0.0, -1.7, .5]] * 4, np.float32),
        "one_big_bin_20000": big_bin,
        # class X selections with more ties than the candidate buffer holds (CTA-wide bisection path of k_fit_big)
        "flat_big_bin_9000": np.c_[5 + rng.random(9000) * 0.5, rng.random(9000) * 0.5, np.full(9000, -1.723), rng.random(9000)].astype(np.float32),
        "two_level_big_bin_9000": np.c_[5 + rng.random(9000) * 0.5, rng.random(9000) * 0.5, np.where(rng.random(9000) < 0.6, -1.75, -1.70), rng.random(9000)].astype(np.float32),
        "wall_in_big_zone0_bin": np.r_[np.c_[4 + rng.random(6000) * 0.05, rng.random(6000) * 0.6, -1.7 + rng.random(6000) * 2.0, rng.random(6000)],
                                       np.c_[3 + rng.random(6000) * 4, rng.random(6000) * 0.6, -1.7 + rng.normal(0, 0.02, 6000), rng.random(6000)]].astype(np.float32),
        "nonfinite": nonfinite,
        "z_equals_flt_min": tomb,
        "rnr_hits": rnr,
        "chunk_boundary_4096": np.c_[5 + rng.random(4096) * 30, rng.random(4096) * 30 - 15, -1.7 + rng.normal(0, 0.05, 4096), rng.random(4096)].astype(np.float32),
        "chunk_boundary_4097": np.c_[5 + rng.random(4097) * 30, rng.random(4097) * 30 - 15, -1.7 + rng.normal(0, 0.05, 4097), rng.random(4097)].astype(np.float32),
    }
    # one call with all cases as separate streams (mixed sizes, including empty), and each alone
    names = list(cases)
    eng = _engine(num_streams=len(names))
    eng.estimate_host([cases[k] for k in names])
    for f, k in enumerate(names):
        orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(cases[k])
        compare_frame(eng, f, orc, cases[k], f"edge/{k}")
    for k in ("empty", "one_point", "nonfinite", "rnr_hits"):
        e1 = _engine()
        e1.estimate_host([cases[k]])
        orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(cases[k])
        compare_frame(e1, 0, orc, cases[k], f"edge-single/{k}")
    # the reference drops a non-RNR point with z == FLT_MIN from both lists (patchworkpp.cpp:591)
    e1 = _engine(); e1.estimate_host([tomb])
    assert e1.num_ground(0) + e1.num_nonground(0) == len(tomb) - 1


def test_n_by_3_input_disables_rnr(kitti):
    a = kitti[0]
    eng = _engine()
    eng.estimate_host([np.ascontiguousarray(a[:, :3])])
    orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(np.ascontiguousarray(a[:, :3]))
    compare_frame(eng, 0, orc, a[:, :3], "Nx3")
    assert not (eng.bin_ids(0) == eng.nbins).any()  # no RNR pseudo-bin


def test_strided_input_layouts(kitti):
    """numpy C order, Eigen-style column-major, and a strided view give the same result without caller-side copies."""
    a = kitti[1]
    eng = _engine()
    eng.estimate_host([a])
    g0 = np.sort(eng.ground_indices(0))
    col = np.asfortranarray(a)
    eng2 = _engine()
    eng2.estimate_host_strided([col.ctypes.data], [a.shape[0]], 4, 1, a.shape[0])
    assert np.array_equal(g0, np.sort(eng2.ground_indices(0)))
    wide = np.zeros((a.shape[0], 6), np.float32); wide[:, 1:5] = a
    eng3 = _engine()
    eng3.estimate_host_strided([wide[:, 1:].ctypes.data], [a.shape[0]], 4, 6, 1)
    assert np.array_equal(g0, np.sort(eng3.ground_indices(0)))


def test_device_resident_input_and_determinism(kitti):
    import torch
    frames = kitti[:3]
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    offs = np.cumsum([0] + [len(f) for f in frames]).astype(np.int64)
    eng = _engine(num_streams=3)
    eng.estimate_device(pts.data_ptr(), offs)
    eng.synchronize()
    res1 = [(eng.ground_indices(f).copy(), eng.nonground_indices(f).copy(), bytes(eng.bin_results(f))) for f in range(3)]
    engh = _engine(num_streams=3)
    engh.estimate_host(frames)
    for f in range(3):
        assert np.array_equal(res1[f][0], engh.ground_indices(f)) and np.array_equal(res1[f][1], engh.nonground_indices(f))
    # bit-reproducible run to run (stable scatter => fixed summation order)
    eng.reset()
    eng.estimate_device(pts.data_ptr(), offs)
    for f in range(3):
        assert np.array_equal(res1[f][0], eng.ground_indices(f)) and np.array_equal(res1[f][1], eng.nonground_indices(f))
        assert res1[f][2] == bytes(eng.bin_results(f))
    # xyz getters return the coordinates of the listed points
    g = eng.ground_indices(1)
    assert np.array_equal(eng.ground_xyz(1), frames[1][g, :3])
    ng = eng.nonground_indices(1)
    assert np.array_equal(eng.nonground_xyz(1), frames[1][ng, :3])


def test_emission_order_is_bin_major(kitti):
    """Bins appear in the reference's emission order; inside a bin indices ascend."""
    a = kitti[0]
    eng = _engine(); eng.estimate_host([a])
    orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(a)
    ids = orc.bin_ids().astype(np.int64)
    eng_g = eng.ground_indices(0)
    for ge, go in ((eng_g, orc.getGroundIndices()), (eng.nonground_indices(0), orc.getNongroundIndices())):
        # same sequence of bins (run-length encoded), same multiset inside each run
        def runs(idx):
            b = ids[idx]
            cut = np.nonzero(np.diff(b))[0] + 1
            return [np.sort(x) for x in np.split(idx, cut)], b[np.r_[0, cut]] if len(idx) else []
        re_, be_ = runs(ge)
        ro_, bo_ = runs(go)
        assert list(be_) == list(bo_)
        assert all(np.array_equal(x, y) for x, y in zip(re_, ro_))
        # inside one bin's run: one ascending piece in the ground list; in the non-ground list at most two
        # (a rejected patch's ground part, then its non-ground part, patchworkpp.cpp:264/272 then :284)
        for x in np.split(ge, np.nonzero(np.diff(ids[ge]))[0] + 1):
            assert int((np.diff(x) <= 0).sum()) <= (0 if ge is eng_g else 1)


def test_python_dropin_module(kitti):
    """The reference's Python call sequence (python/examples/demo_visualize.py:18-39) through pypatchworkpp."""
    import pypatchworkpp
    params = pypatchworkpp.Parameters()
    params.verbose = False
    pw = pypatchworkpp.patchworkpp(params)
    orc = O.Oracle(arith=O.ARITH_CANON64)
    for f in range(2):
        a = kitti[f]
        pw.estimateGround(a)
        orc.estimate(a)
        g, ng = pw.getGroundIndices(), pw.getNongroundIndices()
        assert g.dtype == np.int32 and ng.dtype == np.int32
        assert_sets_equal(orc.getGroundIndices(), orc.getNongroundIndices(), g, ng, len(a), f"pybind/{f}")
        G = pw.getGround()
        assert G.dtype == np.float32 and G.shape == (len(g), 3) and np.array_equal(G, a[g, :3])
        assert pw.getNonground().shape == (len(ng), 3)
        assert pw.getCenters().shape == pw.getNormals().shape == (orc._f("num_patches")(orc._h), 3)
        assert abs(pw.getHeight() - orc.getHeight()) < 1e-9
        assert pw.getTimeTaken() > 0
    # Fortran-ordered and float64 inputs are accepted like pybind11/eigen.h accepts them
    pw2 = pypatchworkpp.patchworkpp(params)
    pw2.estimateGround(np.asfortranarray(kitti[0]).astype(np.float64))
    orc2 = O.Oracle(arith=O.ARITH_CANON64); orc2.estimate(kitti[0])
    assert np.array_equal(np.sort(pw2.getGroundIndices()), np.sort(orc2.getGroundIndices()))


def test_large_batch_properties():
    """BASELINE config-3 shape at reduced count: size-independent invariants on every frame + oracle on a sample."""
    import torch
    import synth
    nf = 128
    pts, offs = synth.make_batch(4242, 0, nf, "kitti64", "cuda")
    eng = _engine(num_streams=nf)
    eng.estimate_device(pts.data_ptr(), offs.numpy())
    eng.synchronize()
    host = pts.cpu().numpy()
    offs = offs.numpy()
    frac = []
    for f in range(nf):
        n = int(offs[f + 1] - offs[f])
        g, ng = eng.ground_indices(f), eng.nonground_indices(f)
        assert len(g) + len(ng) == n
        seen = np.zeros(n, np.int32); seen[g] += 1; seen[ng] += 1
        assert np.all(seen == 1), f"frame {f}: not a partition"
        frac.append(len(g) / n)
    ndeg = 0
    # oracle on a sample plus the frames with the most unusual ground share
    sample = sorted(set(range(0, nf, 8)) | {int(np.argmin(frac)), int(np.argmax(frac))})
    for f in sample:
        a = host[offs[f]:offs[f + 1]]
        orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(a)
        ndeg += compare_frame(eng, f, orc, a, f"large/{f}")
    assert ndeg <= 4
    # idempotence: same input on fresh streams gives identical output
    first = [eng.ground_indices(f).copy() for f in range(0, nf, 16)]
    eng.reset()
    eng.estimate_device(pts.data_ptr(), offs)
    for i, f in enumerate(range(0, nf, 16)):
        assert np.array_equal(first[i], eng.ground_indices(f))


def test_dense_ouster_frame():
    """BASELINE config-5 shape: a ~1M-point frame (bins of tens of thousands of points)."""
    import synth
    a = synth.make_frame(5, 0, "dense1m", "cuda").cpu().numpy()
    assert a.shape[0] > 800_000
    eng = _engine(); eng.estimate_host([a])
    orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(a)
    compare_frame(eng, 0, orc, a, "ouster128")
