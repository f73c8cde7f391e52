"""The CUDA kernels themselves, executed on the CPU, vs the oracle (CANON64).

tests/simt/ compiles csrc/pwpp_kernels.cuh + csrc/pwpp_fit.cuh with g++ against a stand-in for <cuda_runtime.h> that
runs every thread of a CTA as a fiber and implements warp shuffles / ballots / match / reduce and block barriers as
rendezvous (tests/simt/cuda_runtime.h). The launch sequence is the one of launch_range() in csrc/pwpp_capi.cu.
This is the GPU parity suite's bar (tests/test_gpu_parity.py) applied to the kernel code in the GPU-less build
container: bin ids bit-exact, index sets identical, per-patch planes and adaptive state within the CANON64 tolerances.
It cannot see data races or anything about speed; the `-m gpu` suite on a B200 remains the parity gate."""
import numpy as np
import pytest

import oracle_py as O
from helpers import SimtTwin, assert_bins_close, assert_sets_equal, assert_state_close
from param_sets import PARAM_SETS


def _check(orc, tw, a, what, allow_degenerate=False):
    ids = orc.bin_ids()
    assert np.array_equal(ids, tw.bin_ids()), f"{what}: polar bin ids differ"
    g_o, ng_o, g_t, ng_t = orc.getGroundIndices(), orc.getNongroundIndices(), tw.getGroundIndices(), tw.getNongroundIndices()
    assert len(g_o) + len(ng_o) == len(g_t) + len(ng_t), what
    degenerate = orc.bin_min_fit_n() < 3
    if allow_degenerate:
        degenerate |= np.array([orc.bin_results()[b].n < 5 for b in range(orc.nbins)])
    if not degenerate.any():
        assert_sets_equal(g_o, ng_o, g_t, ng_t, len(a), what)
        assert_bins_close(orc.bin_results(), tw.bin_results(), orc.nbins, what)
        assert_state_close(orc.state(), tw.state(), what)
        assert np.abs(orc.getCenters().astype(np.float64) - tw.getCenters()).max(initial=0) <= 1e-6
        assert np.abs(orc.getNormals().astype(np.float64) - tw.getNormals()).max(initial=0) <= 1e-6
        return 0
    keep = ~np.r_[degenerate, np.zeros(3, bool)][ids]
    mo = np.zeros(len(a), bool); mo[g_o] = True
    mt = np.zeros(len(a), bool); mt[g_t] = True
    assert np.array_equal(mo[keep], mt[keep]), f"{what}: labels differ outside degenerate patches"
    return int(degenerate.sum())


def test_fixture_sequence_default(kitti):
    """One stream over the six fixture scans: every kernel incl. the temporal state carried by k_gle."""
    orc, tw = O.Oracle(arith=O.ARITH_CANON64), SimtTwin()
    for f, a in enumerate(kitti):
        orc.estimate(a); tw.estimate(a)
        assert _check(orc, tw, a, f"seq/{f}") == 0
        for r in range(4):
            for w in (0, 1):
                assert np.allclose(tw.history(r, w), orc.history(r, w), rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("pname", ["ros", "no_rvpf_tgr"])
def test_fixture_other_parameter_sets(kitti, pname):
    mk, cols = PARAM_SETS[pname]
    for f in (0, 4):
        a = kitti[f][:, :cols]
        orc, tw = O.Oracle(mk(), O.ARITH_CANON64), SimtTwin(mk())
        orc.estimate(a); tw.estimate(a)
        _check(orc, tw, a, f"{pname}/{f}", allow_degenerate=True)


def test_batched_call_with_mixed_frames(kitti):
    """Several frames in one launch sequence (frame tables, shared work queues, per-frame state), including an empty
    frame, a one-point frame and a synthetic frame; more persistent CTAs than there is work for."""
    import synth
    frames = [kitti[1], np.zeros((0, 4), np.float32), synth.make_frame(5, 1).numpy(), np.array([[5, 0, -1.7, 0.5]], np.float32), kitti[2][:50000]]
    tw = SimtTwin(num_streams=len(frames), persistent_ctas=3)
    tw.estimate_multi(frames)
    for f, a in enumerate(frames):
        orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(a)
        tw.select(f)
        _check(orc, tw, a, f"batch/{f}", allow_degenerate=True)


@pytest.mark.parametrize("opts", [dict(front=0), dict(patch=1), dict(front=0, patch=1), dict(order=1)])
def test_kernel_switches(kitti, opts):
    """The remaining switches (PWPP_FRONT=0: the three stand-alone front-end kernels instead of the cluster kernel;
    PWPP_FIT_PATCH=1: k_fit_patch for the patches above 512 points; the emit split; reference order) give the oracle's result."""
    a = kitti[3]
    orc, tw = O.Oracle(arith=O.ARITH_CANON64), SimtTwin(**opts)
    orc.estimate(a); tw.estimate(a)
    assert _check(orc, tw, a, f"switch/{opts}") == 0


def test_cluster_front_end_is_identical(kitti):
    """k_front_cluster (one cluster of 8 CTAs per frame: TMA-staged tiles, histograms exchanged through distributed shared
    memory) against the three stand-alone kernels: bin ids, index lists and patch records identical, for empty / one-point /
    ragged frames (fewer tiles than CTAs) and under random interleavings of the cluster's CTAs."""
    import synth
    frames = [kitti[1], np.zeros((0, 4), np.float32), synth.make_frame(5, 1).numpy(), np.array([[5, 0, -1.7, 0.5]], np.float32), kitti[2][:50000],
              synth.make_frame(5, 2).numpy()[:4096], kitti[3][:2049]]
    a = SimtTwin(num_streams=len(frames), front=0); a.estimate_multi(frames)
    for seed, mode in ((0, 1), (7, 1), (8, 3)):   # front=3: the 8 x 512-thread shape dense frames get
        b = SimtTwin(num_streams=len(frames), front=mode)
        b.set_sched_seed(seed)
        try:
            b.estimate_multi(frames)
        finally:
            b.set_sched_seed(0)
        for f in range(len(frames)):
            a.select(f); b.select(f)
            assert np.array_equal(a.bin_ids(), b.bin_ids())
            assert np.array_equal(a.getGroundIndices(), b.getGroundIndices()) and np.array_equal(a.getNongroundIndices(), b.getNongroundIndices()), (seed, f)
            assert bytes(a.bin_results()) == bytes(b.bin_results())


def _big_patch_cases():
    rng = np.random.default_rng(3)
    two_level = np.c_[5 + rng.random(9000) * 0.5, rng.random(9000) * 0.5, np.where(rng.random(9000) < 0.6, -1.75, -1.70), rng.random(9000)]
    return {
        "big_bin_20000": np.c_[5 + rng.random(20000) * 0.5, rng.random(20000) * 0.5, -1.7 + rng.normal(0, 0.02, 20000), rng.random(20000)].astype(np.float32),
        "flat_9000": np.c_[5 + rng.random(9000) * 0.5, rng.random(9000) * 0.5, np.full(9000, -1.723), rng.random(9000)].astype(np.float32),
        "two_level_9000": two_level.astype(np.float32),
        "wall_zone0_12000": np.r_[np.c_[4 + rng.random(6000) * 0.05, rng.random(6000) * 0.6, -1.7 + rng.random(6000) * 2.0, rng.random(6000)],
                                  np.c_[3 + rng.random(6000) * 4, rng.random(6000) * 0.6, -1.7 + rng.normal(0, 0.02, 6000), rng.random(6000)]].astype(np.float32),
        "zone2_10000": np.c_[30 + rng.random(10000) * 2, rng.random(10000) * 2, -1.7 + rng.normal(0, 0.03, 10000), rng.random(10000)].astype(np.float32),
    }


@pytest.mark.parametrize("opts", [dict(), dict(order=1)])
def test_big_patches(opts):
    """Class X (more than 8192 points in one patch): k_fit_big, including
    the tie-heavy selections that overflow the candidate buffer and an R-VPF wall removal in zone 0."""
    cases = _big_patch_cases()
    names = list(cases)
    tw = SimtTwin(num_streams=len(names), **opts)
    tw.estimate_multi([cases[k] for k in names])
    for f, k in enumerate(names):
        orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(cases[k])
        tw.select(f)
        assert _check(orc, tw, cases[k], f"big/{k}/{opts}") == 0


def test_patch_class_boundaries():
    """Patch sizes on both sides of every class limit (64 / 512 / 2048 / 4096 / 8192) with both kernel sets."""
    rng = np.random.default_rng(2)
    mk = lambda n: np.c_[5 + rng.random(n) * 0.5, rng.random(n) * 0.5, -1.7 + rng.normal(0, 0.02, n), rng.random(n)].astype(np.float32)  # noqa: E731
    frames = [mk(n) for n in (64, 65, 512, 513, 2048, 2049, 4096, 4097, 8192, 8193)]
    for patch in (0, 1):
        tw = SimtTwin(num_streams=len(frames), patch=patch)
        tw.estimate_multi(frames)
        for f in range(len(frames)):
            tw.select(f)
            orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(frames[f])
            assert _check(orc, tw, frames[f], f"boundary/{patch}/{f}") == 0


def test_dense_frame():
    """BASELINE config-5 shape: one ~1.4M-point frame (27 class-X patches) through every kernel."""
    import synth
    a = synth.make_frame(5, 0, "dense1m").numpy()
    orc, tw = O.Oracle(arith=O.ARITH_CANON64), SimtTwin()
    orc.estimate(a); tw.estimate(a)
    _check(orc, tw, a, "dense1m", allow_degenerate=True)


def test_edge_cases():
    rng = np.random.default_rng(11)
    fill = lambda n, y0: [[5 + 0.01 * i, y0 + rng.random() * 0.3, -1.7 + rng.normal(0, 0.01), .5] for i in range(n)]  # noqa: E731
    cases = {
        "nine_in_one_bin": np.c_[5 + rng.random(9) * 0.1, rng.random(9) * 0.1, -1.7 + rng.random(9) * 0.01, rng.random(9)].astype(np.float32),
        "ten_in_one_bin": np.c_[5 + rng.random(10) * 0.1, rng.random(10) * 0.1, -1.7 + rng.random(10) * 0.01, rng.random(10)].astype(np.float32),
        "all_out_of_range": np.c_[rng.random((50, 2)) * 1.0, rng.random((50, 2))].astype(np.float32),
        "flat_plane": np.c_[(rng.random((5000, 2)) - 0.5) * 60, np.full(5000, -1.723), rng.random(5000)].astype(np.float32),
        "one_big_bin_20000": np.c_[5 + rng.random(20000) * 0.5, rng.random(20000) * 0.5, -1.7 + rng.normal(0, 0.02, 20000), rng.random(20000)].astype(np.float32),
        "one_bin_6000": np.c_[5 + rng.random(6000) * 0.5, rng.random(6000) * 0.5, -1.7 + rng.normal(0, 0.02, 6000), rng.random(6000)].astype(np.float32),
        "nonfinite": np.array([[5, 1, np.nan, .5], [np.nan, 1, -1.7, .5], [5, np.inf, -1.7, .5], [6, 1, -np.inf, .5], [6, 1, np.inf, .01], [7, 2, -1.7, np.nan]] + fill(30, 1.0), np.float32),
        "z_equals_flt_min": np.array([[5, 1, np.finfo(np.float32).tiny, .5]] + fill(15, 1.2), np.float32),
        "rnr_hits": np.array([[4, 0, -3.0, 0.05], [4, 0.1, -3.0, 0.5], [4, 0.2, -2.4, 0.05], [40, 0.2, -3.0, 0.05]] + fill(12, 1.0), np.float32),
        "chunk_boundary_4097": np.c_[5 + rng.random(4097) * 30, rng.random(4097) * 30 - 15, -1.7 + rng.normal(0, 0.05, 4097), rng.random(4097)].astype(np.float32),
        "vertical_wall_zone0": np.r_[np.c_[4 + rng.random(3000) * 0.05, rng.random(3000) * 1.5, -1.7 + rng.random(3000) * 2.0, rng.random(3000)],
                                     np.c_[3 + rng.random(3000) * 6, rng.random(3000) * 1.5, -1.7 + rng.normal(0, 0.02, 3000), rng.random(3000)]].astype(np.float32),
    }
    names = list(cases)
    for opts in (dict(), dict(patch=1, front=0)):
        tw = SimtTwin(num_streams=len(names), **opts)
        tw.estimate_multi([cases[k] for k in names])
        for f, k in enumerate(names):
            orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(cases[k])
            tw.select(f)
            _check(orc, tw, cases[k], f"edge/{k}/{opts}", allow_degenerate=True)
    tw.select(names.index("z_equals_flt_min"))
    assert len(tw.getGroundIndices()) + len(tw.getNongroundIndices()) == len(cases["z_equals_flt_min"]) - 1   # patchworkpp.cpp:591


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15, 16, 17, 18])
def test_random_parameter_sets(kitti, seed):
    """Random supported parameter sets (bin layouts, num_iter / num_lpr / num_min_pts, thresholds, range, switches on and
    off, N x 3 input) and random kernel-variant switches, a three-frame sequence each: the kernels on the SIMT twin vs
    the oracle. Patches the algorithm leaves numerically undefined (fits to < 3 points) are excluded as everywhere."""
    import synth
    from pwpp_ctypes import default_params
    rng = np.random.default_rng(seed)
    p = default_params()
    p.enable_RNR, p.enable_RVPF, p.enable_TGR = (int(rng.random() < 0.7) for _ in range(3))
    p.num_iter = int(rng.integers(1, 6)); p.num_lpr = int(rng.choice([1, 5, 20, 32, 33, 64])); p.num_min_pts = int(rng.choice([0, 1, 3, 5, 10, 30, 200]))
    p.num_rings_of_interest = int(rng.integers(0, 5))
    p.max_flatness_storage = int(rng.choice([3, 40, 1000])); p.max_elevation_storage = int(rng.choice([2, 50, 1000]))
    p.sensor_height = float(rng.uniform(1.4, 2.1))
    p.th_seeds = float(rng.choice([0.05, 0.125, 0.3, 0.5])); p.th_seeds_v = float(rng.choice([0.05, 0.25, 0.4]))
    p.th_dist = float(rng.choice([0.05, 0.125, 0.3])); p.th_dist_v = float(rng.choice([0.05, 0.1, 0.9]))
    p.min_range = float(rng.choice([0.0, 0.5, 2.7, 5.0])); p.max_range = float(rng.choice([10.0, 40.0, 80.0, 120.0, 250.0, 300.0]))
    p.uprightness_thr = float(rng.choice([0.101, 0.5, 0.707, 0.95]))
    p.num_sectors_each_zone[:] = [int(x) for x in rng.choice([1, 4, 8, 16, 32, 54, 64, 128, 200], 4)]
    p.num_rings_each_zone[:] = [int(x) for x in rng.integers(1, 9, 4)]
    rng.choice([1, 3, 8]); rng.integers(0, 2)   # (two draws kept so that the parameter sets below stay the ones of the earlier runs)
    opts = dict(front=int(rng.integers(0, 2)), patch=int(rng.integers(0, 2)))
    cols = 4 if rng.random() < 0.8 else 3
    pool = [kitti[0], kitti[4], synth.make_frame(7, 0).numpy()]
    orc, tw = O.Oracle(p, O.ARITH_CANON64), SimtTwin(p, **opts)
    for t in range(3):
        a = pool[int(rng.integers(0, len(pool)))][:, :cols]
        orc.estimate(a); tw.estimate(a)
        if _check(orc, tw, a, f"random/{seed}/{t}/{opts}", allow_degenerate=True):
            break   # states may legitimately diverge after a numerically undefined patch


@pytest.mark.parametrize("seed", [101, 102, 103, 104, 105, 106])
def test_mutated_inputs(kitti, seed):
    """Inputs a real pipeline can produce but the fixtures do not contain: millimetre-quantised coordinates (exact z ties),
    duplicated points, non-finite values, z == FLT_MIN, intensities on the RNR threshold, rescaled and shuffled scans.
    Bin ids must always be bit-exact and every point emitted once; labels are compared in all patches that are
    well-conditioned for the reference itself (its fp32 and the canonical double arithmetic agree there, and no plane
    was fitted to fewer than 4 points anywhere in the frame)."""
    import synth
    rng = np.random.default_rng(seed)
    pool = [kitti[0], kitti[2], synth.make_frame(7, 1).numpy()]
    a = pool[int(rng.integers(0, len(pool)))].copy()
    a = a[: int(rng.integers(20000, len(a)))]
    if rng.random() < 0.7:
        q = float(rng.choice([0.001, 0.002, 0.01])); a[:, 2] = np.round(a[:, 2] / q) * q
    if rng.random() < 0.4:
        q = float(rng.choice([0.001, 0.002, 0.01])); a[:, :2] = np.round(a[:, :2] / q) * q
    if rng.random() < 0.5:
        a = np.concatenate([a, a[rng.integers(0, len(a), int(rng.integers(1, 2000)))]])
    if rng.random() < 0.6:
        k = int(rng.integers(1, 200)); idx = rng.integers(0, len(a), k)
        vals = np.array([np.nan, np.inf, -np.inf, np.finfo(np.float32).tiny, 0.0, -0.0, 80.0, 2.7, -2.7], np.float32)
        a[idx, rng.integers(0, 4, k)] = vals[rng.integers(0, len(vals), k)]
    if rng.random() < 0.5:
        k = int(rng.integers(1, 500)); idx = rng.integers(0, len(a), k)
        a[idx, 3] = np.array([0.2, 0.19999999, 0.20000002, 0.0, 1.0], np.float32)[rng.integers(0, 5, k)]
        a[idx, 2] = np.float32(-2.6) + rng.normal(0, 0.2, k).astype(np.float32)
    if rng.random() < 0.3:
        a[:, :3] *= np.float32(rng.choice([0.5, 2.0]))
    if rng.random() < 0.3:
        a = a[rng.permutation(len(a))]
    opts = dict(front=int(rng.integers(0, 2)), patch=int(rng.integers(0, 2)))
    orc, ref, tw = O.Oracle(arith=O.ARITH_CANON64), O.Oracle(arith=O.ARITH_REF32), SimtTwin(**opts)
    orc.estimate(a); ref.estimate(a); tw.estimate(a)
    ids = orc.bin_ids()
    assert np.array_equal(ids, tw.bin_ids()), f"{opts}: bin ids differ"
    g_o, ng_o, g_t, ng_t = orc.getGroundIndices(), orc.getNongroundIndices(), tw.getGroundIndices(), tw.getNongroundIndices()
    assert len(g_o) + len(ng_o) == len(g_t) + len(ng_t)
    allidx = np.concatenate([g_t, ng_t])
    assert len(np.unique(allidx)) == len(allidx)
    if (orc.bin_min_fit_n() < 4).any():
        return   # a rank-deficient fit somewhere: its patch and, through the ring statistics of TGR, its neighbours are undefined
    mo = np.zeros(len(a), bool); mo[g_o] = True
    mr = np.zeros(len(a), bool); mr[ref.getGroundIndices()] = True
    mt = np.zeros(len(a), bool); mt[g_t] = True
    illcond = np.zeros(orc.nbins + 3, bool)
    illcond[np.unique(ids[mo != mr])] = True
    keep = ~illcond[ids]
    assert np.array_equal(mo[keep], mt[keep]), f"{opts}: {int((mo[keep] != mt[keep]).sum())} labels differ in well-conditioned patches"


def test_history_bound_drops_the_oldest_samples():
    """The defined deviation of include/pwpp.h: while ring 0 holds <= 1 flatness samples the reference never trims the flatness
    histories of rings 1..3 (the `break` of S:363-364) and they grow without bound; the kernels keep the NEWEST hist_cap =
    max(max_*_storage) + 4 * (max sectors) + 64 samples. A stream whose scans leave ring 0 empty for 25 frames: until the bound
    is reached everything equals the oracle; past it the labels and the thresholds still do (the thresholds are frozen while the
    break holds), the ring-1 history is the oracle's newest hist_cap samples; when ring 0 finally gets its samples the flatness
    threshold of ring 1 is mean + stdev over exactly those samples (the reference would average over its whole history)."""
    from pwpp_ctypes import default_params
    def mk():
        p = default_params()
        p.max_flatness_storage = 40; p.max_elevation_storage = 40
        return p
    hcap = 40 + 4 * 54 + 64
    rng = np.random.default_rng(11)

    def ring_points(r0, r1, nsec, per_sector):
        ang = (np.arange(nsec * per_sector) // per_sector + rng.random(nsec * per_sector) * 0.9 + 0.05) * (2 * np.pi / nsec)
        rad = r0 + rng.random(len(ang)) * (r1 - r0)
        return np.c_[rad * np.cos(ang), rad * np.sin(ang), -1.723 + rng.normal(0, 0.01, len(ang)), rng.random(len(ang))].astype(np.float32)

    orc, tw = O.Oracle(mk(), O.ARITH_CANON64), SimtTwin(mk())
    nblocked = 25
    for f in range(nblocked):   # ring 1 of zone 0 (7.6 .. 12.3 m) and ring 0 of zone 1 only: ring 0 of zone 0 stays empty
        a = np.r_[ring_points(7.7, 12.2, 16, 24), ring_points(12.5, 14.6, 32, 16)]
        orc.estimate(a); tw.estimate(a)
        assert _check_labels_only(orc, tw, a, f"blocked/{f}")
        so, st = orc.state(), tw.state()
        assert so.n_flatness[0] == 0 and st.n_flatness[0] == 0
        assert st.n_flatness[1] == min(so.n_flatness[1], hcap) and st.n_flatness[2] == min(so.n_flatness[2], hcap)
        for r in (1, 2):
            ho, ht = orc.history(r, 1), tw.history(r, 1)
            assert np.allclose(ht, ho[-len(ht):], rtol=1e-6, atol=1e-9), f"ring {r}: not the newest samples"
        assert list(so.flatness_thr) == list(st.flatness_thr)   # frozen while the break holds
    assert orc.state().n_flatness[1] > hcap   # the scenario did pass the bound
    held = tw.history(1, 1).copy()                      # what the bounded row holds: the oracle's newest hist_cap samples (asserted above)
    assert len(held) == hcap
    a = np.r_[ring_points(3.0, 7.4, 16, 24), ring_points(7.7, 12.2, 16, 24)]   # ring 0 gets 16 patches: > 1 samples, the break is gone
    orc.estimate(a); tw.estimate(a)
    so, st = orc.state(), tw.state()
    assert st.n_flatness[0] > 1 and st.n_flatness[1] == 40 and so.n_flatness[1] == 40     # every ring updated and trimmed (S:372-373)
    assert np.allclose(tw.history(1, 1), orc.history(1, 1), rtol=1e-6, atol=1e-9)        # the newest 40 agree again
    # the threshold of ring 1 was computed over the newest hist_cap of (held samples + this frame's k new ones), k = 0..16 unknown here
    post = orc.history(1, 1)
    def thr(v):   # calc_mean_stdev + S:368 in the same sequential arithmetic
        m = 0.0
        for x in v: m += x
        m /= len(v)
        q = 0.0
        for x in v: q += (x - m) * (x - m)
        return m + (q / (len(v) - 1)) ** 0.5
    cands = [thr(list(np.r_[held, post[len(post) - k:]][-hcap:])) for k in range(0, 17)]
    assert min(abs(st.flatness_thr[1] - c) for c in cands) <= 1e-12 * max(1.0, abs(st.flatness_thr[1])), (st.flatness_thr[1], cands[:3])
    # ... and it is NOT what the unbounded reference history gives (the documented deviation), while ring 0 — never bounded — agrees
    assert abs(st.flatness_thr[1] - so.flatness_thr[1]) > 0 and abs(st.flatness_thr[0] - so.flatness_thr[0]) <= 1e-9


def _check_labels_only(orc, tw, a, what):
    assert np.array_equal(orc.bin_ids(), tw.bin_ids()), f"{what}: polar bin ids differ"
    assert_sets_equal(orc.getGroundIndices(), orc.getNongroundIndices(), tw.getGroundIndices(), tw.getNongroundIndices(), len(a), what)
    return True


def test_emit_tile_boundaries(kitti):
    """k_emit cuts a frame's bin-sorted order into 1024-position tiles per warp (8192 per CTA): frame sizes on and around those
    limits, in one batched call (the grid is sized for the largest frame: the others end inside it)."""
    sizes = [1, 1023, 1024, 1025, 2049, 8191, 8192, 8193, 16385, 40000]
    frames = [np.ascontiguousarray(kitti[(i % 3)][i * 7:i * 7 + 2 * n:2][:n]) for i, n in enumerate(sizes)]
    tw = SimtTwin(num_streams=len(frames))
    tw.estimate_multi(frames)
    for f, a in enumerate(frames):
        orc = O.Oracle(arith=O.ARITH_CANON64)
        orc.estimate(a)
        tw.select(f)
        assert len(a) == sizes[f]
        _check(orc, tw, a, f"emit tiles / {sizes[f]} points", allow_degenerate=True)
