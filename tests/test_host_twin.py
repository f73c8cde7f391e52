"""CPU twin of the CUDA pipeline vs the oracle (CANON64).

tests/host_twin.cu runs the product's own host+device math (csrc/pwpp_math.cuh) and tests/gle_sequential.cuh sequentially on
the CPU with the same algorithmic restructuring as the kernels. Agreement with the oracle validates, without a GPU:
the fp32-filtered polar binning (bit-exact bin ids), the 3x3 Jacobi SVD, the shifted one-pass covariance, the
sort-free LPR selection, the R-VPF 'alive' test by stored planes, A-GLE/TGR/threshold logic and the output layout."""
import numpy as np
import pytest

import oracle_py as O
from helpers import SimtTwin, Twin, assert_bins_close, assert_sets_equal, assert_state_close
from param_sets import PARAM_SETS


def _compare(orc, tw, a, what, nondegenerate_only=False):
    orc.estimate(a); tw.estimate(a)
    assert np.array_equal(orc.bin_ids(), tw.bin_ids()), f"{what}: bin ids differ"
    assert tw.fast_mismatches() == 0, f"{what}: filtered binning disagreed with the exact path"
    if nondegenerate_only:
        ids = orc.bin_ids()
        bad = orc.bin_min_fit_n() < 3
        bad |= np.array([orc.bin_results()[b].n < 5 for b in range(orc.nbins)])
        bad = np.r_[bad, np.zeros(3, bool)]
        keep = ~bad[ids]
        mo = np.zeros(len(a), bool); mo[orc.getGroundIndices()] = True
        mt = np.zeros(len(a), bool); mt[tw.getGroundIndices()] = True
        assert np.array_equal(mo[keep], mt[keep]), f"{what}: labels differ outside degenerate patches"
        return
    assert_sets_equal(orc.getGroundIndices(), orc.getNongroundIndices(), tw.getGroundIndices(), tw.getNongroundIndices(), len(a), what)
    assert_bins_close(orc.bin_results(), tw.bin_results(), orc.nbins, what)
    assert_state_close(orc.state(), tw.state(), what)
    assert np.array_equal(orc.getCenters(), tw.getCenters()) or np.abs(orc.getCenters() - tw.getCenters()).max() < 1e-6


@pytest.mark.parametrize("pname", ["default", "no_rvpf_tgr"])
def test_twin_matches_oracle_on_fixture_sequence(kitti, pname):
    mk, cols = PARAM_SETS[pname]
    orc, tw = O.Oracle(mk(), O.ARITH_CANON64), Twin(mk())
    for f, a in enumerate(kitti):
        deg = bool((O.Oracle(mk(), O.ARITH_CANON64)).nbins) and pname != "default"
        _compare(orc, tw, a[:, :cols], f"{pname}/seq/{f}", nondegenerate_only=deg)


def test_twin_matches_oracle_ros_nondegenerate(kitti):
    mk, cols = PARAM_SETS["ros"]
    for f in (0, 3):
        _compare(O.Oracle(mk(), O.ARITH_CANON64), Twin(mk()), kitti[f][:, :cols], f"ros/fresh/{f}", nondegenerate_only=True)


def test_twin_matches_oracle_on_synthetic():
    import synth
    orc, tw = O.Oracle(arith=O.ARITH_CANON64), Twin()
    for f in range(5):
        _compare(orc, tw, synth.make_frame(99, f).numpy(), f"synthetic/{f}")


def test_binning_filter_exact_on_adversarial_points():
    """Points on and next to every ring / sector / range boundary: the fp32 filter must fall back to the exact path."""
    orc, tw = O.Oracle(arith=O.ARITH_CANON64), Twin()
    rs = [2.7, 12.3625, 22.025, 41.35, 80.0, 7.53125, 14.778125, 31.6875, 60.675]
    pts = []
    for r in rs:
        for dr in (0.0, 1e-7, -1e-7, 1e-5, -1e-5, 3e-4, -3e-4):
            for k in range(0, 108):
                th = k * (2 * np.pi / 108)
                for dth in (0.0, 1e-8, -1e-8, 1e-6, -1e-6):
                    pts.append([(r + dr) * np.cos(th + dth), (r + dr) * np.sin(th + dth), -1.7, 0.5])
    a = np.array(pts, np.float32)
    orc.estimate(a); tw.estimate(a)
    assert np.array_equal(orc.bin_ids(), tw.bin_ids())
    assert tw.fast_mismatches() == 0


def _random_spd(rng, n):
    """n symmetric PSD 3x3 matrices with prescribed spectra: generic, clustered low pair, clustered high pair, rank 1/2, scaled."""
    q, _ = np.linalg.qr(rng.normal(size=(n, 3, 3)))
    lam = np.sort(rng.uniform(0.0, 1.0, size=(n, 3)), axis=1)
    kind = rng.integers(0, 6, size=n)
    lam[kind == 1, 1] = lam[kind == 1, 0] * (1 + 10.0 ** rng.uniform(-12, -1, size=(kind == 1).sum()))     # low pair clustered
    lam[kind == 2, 1] = lam[kind == 2, 2] * (1 - 10.0 ** rng.uniform(-12, -1, size=(kind == 2).sum()))     # high pair clustered
    lam[kind == 3, 0] = 0.0                                                                               # rank 2 (flat patch)
    lam[kind == 4, :2] *= 10.0 ** rng.uniform(-10, -2, size=((kind == 4).sum(), 1))                          # nearly a line
    lam = np.sort(lam, axis=1)
    scale = 10.0 ** rng.uniform(-8, 6, size=(n, 1))
    lam = lam * scale
    a = np.einsum("nij,nj,nkj->nik", q, lam, q)
    a = 0.5 * (a + a.transpose(0, 2, 1))
    return a, lam


def test_closed_form_eigensolver_against_lapack_and_jacobi():
    """The device's 3x3 solver (csrc/pwpp_math.cuh sym_eig3: Newton root of the characteristic cubic + 2x2 deflation)
    vs numpy's eigh and the Jacobi SVD the oracle uses. Tolerances are the conditioning limit of the problem:
    singular values to 16 eps * lambda_max, the normal to 16 eps * lambda_max / (lambda_mid - lambda_min)."""
    import ctypes as C
    from helpers import HERE
    import os
    lib = C.CDLL(os.path.join(HERE, "_build", "libpwpp_twin.so"))
    for f in (lib.twin_sym_eig3, lib.twin_jacobi_svd3): f.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
    rng = np.random.default_rng(5)
    n = 200_000
    a, _ = _random_spd(rng, n)
    cov = np.ascontiguousarray(np.stack([a[:, 0, 0], a[:, 0, 1], a[:, 0, 2], a[:, 1, 1], a[:, 1, 2], a[:, 2, 2]], axis=1))
    out = np.empty((n, 6)); jac = np.empty((n, 6))
    lib.twin_sym_eig3(cov.ctypes.data, n, out.ctypes.data)
    lib.twin_jacobi_svd3(cov.ctypes.data, n, jac.ctypes.data)
    w, v = np.linalg.eigh(a)                      # ascending
    eps = np.finfo(float).eps
    lmax = np.abs(w).max(axis=1)
    sv_ref = np.sort(np.abs(w), axis=1)[:, ::-1]
    assert (np.abs(out[:, :3] - sv_ref).max(axis=1) <= 16 * eps * lmax).all()
    assert (np.abs(jac[:, :3] - sv_ref).max(axis=1) <= 64 * eps * lmax).all()
    vec = out[:, 3:]
    assert np.abs(np.linalg.norm(vec, axis=1) - 1).max() < 8 * eps
    # residual of the eigen-pair: |A v - l v| <= 16 eps lmax — holds regardless of clustering
    lmin = out[:, 2]
    res = np.linalg.norm(np.einsum("nij,nj->ni", a, vec) - lmin[:, None] * vec, axis=1)
    assert (res <= 16 * eps * lmax).all(), float((res / (eps * lmax)).max())
    gap = w[:, 1] - w[:, 0]
    ok = gap > 1e3 * eps * lmax
    ref = v[:, :, 0]
    sgn = np.sign(np.einsum("ni,ni->n", ref, vec)); sgn[sgn == 0] = 1
    err = np.linalg.norm(vec - sgn[:, None] * ref, axis=1)
    assert (err[ok] <= 16 * eps * lmax[ok] / gap[ok]).all(), float((err[ok] * gap[ok] / (eps * lmax[ok])).max())
    # degenerate inputs keep the contract of jacobi_svd3
    special = np.array([[0, 0, 0, 0, 0, 0], [2, 0, 0, 2, 0, 2], [3, 0, 0, 1, 0, 2], [np.nan, 0, 0, 1, 0, 1], [np.inf, 0, 0, 1, 0, 1],
                        [1e-300, 0, 0, 1e-300, 1e-301, 1e-300], [1e300, 1e299, 0, 1e300, 0, 1e300]], dtype=float)
    o2 = np.empty((len(special), 6)); lib.twin_sym_eig3(special.ctypes.data, len(special), o2.ctypes.data)
    assert np.array_equal(o2[0], [0, 0, 0, 0, 0, 1])
    assert np.allclose(o2[1, :3], 2) and np.allclose(o2[2, :3], [3, 2, 1]) and np.allclose(np.abs(o2[2, 3:]), [0, 1, 0])
    assert np.isnan(o2[3, :3]).all() and np.isnan(o2[4, :3]).all() and np.array_equal(o2[3, 3:], [0, 0, 1])
    assert np.isfinite(o2[5:]).all() and np.allclose(o2[6, :3] / 1e300, [1.1, 1.0, 0.9])


@pytest.mark.parametrize("min_range,max_range,rings", [(5.0, 40.0, [5, 1, 2, 3]), (2.7, 80.0, [2, 4, 4, 4]), (1.0, 30.0, [2, 2, 2, 8]), (0.5, 120.0, [1, 1, 1, 1]),
                                                       (5.0, 17.0, [1, 1, 1, 4]), (2.0, 26.0, [2, 2, 2, 5])])
def test_binning_filter_range_limits_for_other_geometries(min_range, max_range, rings):
    """Points a hair inside / outside min_range and max_range and around every ring boundary, for ring widths from 0.3 m to
    60 m: the fp32 filter (used only when every ring is at least 1.5 m wide) and the exact path must both agree with the
    oracle. Regression for a point 1.9e-4 m below min_range that the filter binned with rings 0.875 m wide."""
    from pwpp_ctypes import default_params
    p = default_params()
    p.min_range, p.max_range = min_range, max_range
    p.num_rings_each_zone[:] = rings
    p.num_sectors_each_zone[:] = [32, 54, 32, 32]
    orc, tw = O.Oracle(p, O.ARITH_CANON64), Twin(p)
    z = [min_range, (7 * min_range + max_range) / 8, (3 * min_range + max_range) / 4, (min_range + max_range) / 2, max_range]
    radii = list(z)
    for k in range(4):
        radii += [z[k] + (z[k + 1] - z[k]) * i / rings[k] for i in range(1, rings[k])]
    pts = []
    for r in radii:
        for dr in (0.0, 1e-7, -1e-7, 1e-5, -1e-5, 1.5e-4, -1.5e-4, 1.9e-4, -1.9e-4, 2.1e-4, -2.1e-4, 2.6e-4, -2.6e-4, 4e-4, -4e-4, 1e-3, -1e-3):
            for k in range(0, 54, 3):
                th = k * (2 * np.pi / 54) + 0.013
                pts.append([(r + dr) * np.cos(th), (r + dr) * np.sin(th), -1.7, 0.5])
    a = np.array(pts, np.float32)
    orc.estimate(a); tw.estimate(a)
    assert np.array_equal(orc.bin_ids(), tw.bin_ids()), f"{int((orc.bin_ids() != tw.bin_ids()).sum())} bin ids differ"
    assert tw.fast_mismatches() == 0
    stw = SimtTwin(p)
    stw.estimate(a)
    assert np.array_equal(orc.bin_ids(), stw.bin_ids())
