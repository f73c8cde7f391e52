"""CPU twin of the CUDA pipeline vs the oracle (CANON64).

tests/host_twin.cu runs the product's own host+device code (csrc/pwpp_math.cuh, csrc/pwpp_gle.cuh) sequentially on
the CPU with the same algorithmic restructuring as the kernels. Agreement with the oracle validates, without a GPU:
the fp32-filtered polar binning (bit-exact bin ids), the 3x3 Jacobi SVD, the shifted one-pass covariance, the
sort-free LPR selection, the R-VPF 'alive' test by stored planes, A-GLE/TGR/threshold logic and the output layout."""
import numpy as np
import pytest

import oracle_py as O
from helpers import Twin, assert_bins_close, assert_sets_equal, assert_state_close
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
