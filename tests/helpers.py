"""Shared comparison helpers for the parity tests."""
import ctypes as C
import os

import numpy as np

import oracle_py as O
from oracle_py import _Base
from pwpp_ctypes import PwppBinResult, PwppParams, default_params

HERE = os.path.dirname(os.path.abspath(__file__))


class Twin(_Base):
    """tests/host_twin.cu: the product's host+device math run sequentially on the CPU."""

    def __init__(self, params=None):
        lib = C.CDLL(os.path.join(HERE, "_build", "libpwpp_twin.so"))
        self._bind(lib, "twin_")
        lib.twin_create.argtypes = [C.POINTER(PwppParams)]; lib.twin_create.restype = C.c_void_p
        lib.twin_bin_ids.argtypes = [C.c_void_p, C.c_void_p]
        lib.twin_bin_results.argtypes = [C.c_void_p, C.c_void_p]
        lib.twin_num_bins.argtypes = [C.c_void_p]
        lib.twin_uses_fast_binning.argtypes = [C.c_void_p]
        lib.twin_fast_mismatches.argtypes = [C.c_void_p]; lib.twin_fast_mismatches.restype = C.c_longlong
        self._lib = lib
        self.params = params if params is not None else default_params()
        self._h = lib.twin_create(C.byref(self.params))
        self.nbins = lib.twin_num_bins(self._h)

    def bin_ids(self):
        out = np.empty(self._n, dtype=np.uint16)
        self._lib.twin_bin_ids(self._h, out.ctypes.data)
        return out

    def bin_results(self):
        arr = (PwppBinResult * self.nbins)()
        self._lib.twin_bin_results(self._h, C.byref(arr))
        return arr

    def fast_mismatches(self):
        return self._lib.twin_fast_mismatches(self._h)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else (a.view(np.uint64) if a.dtype == np.float64 else a)


def assert_bit_identical(a, b, what=""):
    """Everything the reference surface exposes, bit for bit, including emission order."""
    for name in ("getGroundIndices", "getNongroundIndices", "getGround", "getNonground", "getCenters", "getNormals"):
        x, y = getattr(a, name)(), getattr(b, name)()
        assert x.shape == y.shape, f"{what} {name}: shape {x.shape} vs {y.shape}"
        assert np.array_equal(bits(x), bits(y)), f"{what} {name}: values differ"
    sa, sb = a.state(), b.state()
    assert bits(np.float64(sa.sensor_height)) == bits(np.float64(sb.sensor_height)), f"{what} sensor_height"
    for fld in ("elevation_thr", "flatness_thr"):
        assert np.array_equal(bits(np.array(getattr(sa, fld))), bits(np.array(getattr(sb, fld)))), f"{what} {fld}"
    assert list(sa.n_elevation) == list(sb.n_elevation) and list(sa.n_flatness) == list(sb.n_flatness), f"{what} history sizes"
    for r in range(4):
        for w in (0, 1):
            assert np.array_equal(bits(a.history(r, w)), bits(b.history(r, w))), f"{what} history ring {r} kind {w}"


# stated floating-point tolerances of the CANON64 contract
TOL_NORMAL = 1e-9       # |delta| per component of a unit normal, device/twin vs oracle (both double)
TOL_MEAN = 1e-9         # metres
TOL_SV_REL = 1e-6       # relative to the largest singular value of the patch
TOL_STATE = 1e-9        # sensor height / elevation thresholds (m); flatness thresholds relative 1e-6


def assert_sets_equal(g_a, ng_a, g_b, ng_b, n, what=""):
    """Index SETS identical and a partition of [0, n)."""
    ga, gb = np.sort(np.asarray(g_a)), np.sort(np.asarray(g_b))
    na, nbb = np.sort(np.asarray(ng_a)), np.sort(np.asarray(ng_b))
    assert ga.shape == gb.shape and np.array_equal(ga, gb), f"{what}: ground index sets differ ({len(ga)} vs {len(gb)}; sym diff {len(np.setxor1d(ga, gb))})"
    assert na.shape == nbb.shape and np.array_equal(na, nbb), f"{what}: non-ground index sets differ"
    allidx = np.concatenate([ga, na])
    assert len(np.unique(allidx)) == len(allidx), f"{what}: an index appears twice"


def assert_bins_close(ba, bb, nbins, what=""):
    for b in range(nbins):
        x, y = ba[b], bb[b]
        assert (x.n, x.n_ground, x.verdict, x.fitted) == (y.n, y.n_ground, y.verdict, y.fitted), \
            f"{what} bin {b}: (n,n_ground,verdict,fitted) {(x.n, x.n_ground, x.verdict, x.fitted)} vs {(y.n, y.n_ground, y.verdict, y.fitted)}"
        if not x.fitted or x.n == 0:
            continue
        smax = max(abs(x.sv[0]), 1e-300)
        for k in range(3):
            if np.isnan(x.sv[k]) or np.isnan(y.sv[k]):
                assert np.isnan(x.sv[k]) and np.isnan(y.sv[k]), f"{what} bin {b}: NaN singular value mismatch"
                continue
            assert abs(x.normal[k] - y.normal[k]) <= TOL_NORMAL, f"{what} bin {b} normal[{k}] {x.normal[k]} vs {y.normal[k]}"
            assert abs(x.mean[k] - y.mean[k]) <= TOL_MEAN, f"{what} bin {b} mean[{k}] {x.mean[k]} vs {y.mean[k]}"
            assert abs(x.sv[k] - y.sv[k]) <= TOL_SV_REL * smax, f"{what} bin {b} sv[{k}] {x.sv[k]} vs {y.sv[k]}"


def assert_state_close(sa, sb, what=""):
    assert abs(sa.sensor_height - sb.sensor_height) <= TOL_STATE, f"{what} sensor_height {sa.sensor_height} vs {sb.sensor_height}"
    for i in range(4):
        assert abs(sa.elevation_thr[i] - sb.elevation_thr[i]) <= TOL_STATE, f"{what} elevation_thr[{i}]"
        assert abs(sa.flatness_thr[i] - sb.flatness_thr[i]) <= 1e-6 * max(abs(sb.flatness_thr[i]), 1e-12), f"{what} flatness_thr[{i}]"
    assert list(sa.n_elevation) == list(sb.n_elevation) and list(sa.n_flatness) == list(sb.n_flatness), f"{what} history sizes"


class SimtTwin(_Base):
    """tests/simt/simt_twin.cpp: the CUDA kernels themselves, executed on the CPU by the fiber-based SIMT stand-in."""

    def __init__(self, params=None, num_streams=1, **options):
        lib = C.CDLL(os.path.join(HERE, "_build", "libpwpp_simt.so"))
        self._bind(lib, "simt_")
        lib.simt_create.argtypes = [C.POINTER(PwppParams), C.c_int]; lib.simt_create.restype = C.c_void_p
        lib.simt_bin_ids.argtypes = [C.c_void_p, C.c_void_p]
        lib.simt_bin_results.argtypes = [C.c_void_p, C.c_void_p]
        lib.simt_num_bins.argtypes = [C.c_void_p]
        lib.simt_select.argtypes = [C.c_void_p, C.c_int]
        lib.simt_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.simt_queue_sizes.argtypes = [C.c_void_p]; lib.simt_queue_sizes.restype = C.c_char_p
        lib.simt_estimate_multi.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.simt_set_sched_seed.argtypes = [C.c_ulonglong]
        self._lib = lib
        self.params = params if params is not None else default_params()
        self._h = lib.simt_create(C.byref(self.params), num_streams)
        self.nbins = lib.simt_num_bins(self._h)
        self._ns = [0] * num_streams
        for k, v in options.items():
            assert lib.simt_set_option(self._h, k.encode(), int(v)) == 0, k

    def estimate(self, pts):
        super().estimate(pts)
        self._ns[0] = self._n

    def estimate_multi(self, frames):
        frames = [np.ascontiguousarray(a, dtype=np.float32) for a in frames]
        cols = frames[0].shape[1]
        ptrs = (C.c_void_p * len(frames))(*[a.ctypes.data for a in frames])
        ns = (C.c_int64 * len(frames))(*[a.shape[0] for a in frames])
        self._lib.simt_estimate_multi(self._h, len(frames), ptrs, ns, cols)
        self._ns[:len(frames)] = [a.shape[0] for a in frames]
        self.select(0)

    def select(self, f):
        self._lib.simt_select(self._h, f)
        self._n = self._ns[f]

    def set_sched_seed(self, seed: int):
        """Non-zero: concurrently live CTAs are interleaved at random (per seed) instead of round-robin. Process-wide."""
        self._lib.simt_set_sched_seed(seed)

    def queue_sizes(self):
        return self._lib.simt_queue_sizes(self._h).decode()

    def bin_ids(self):
        out = np.empty(self._n, dtype=np.uint16)
        self._lib.simt_bin_ids(self._h, out.ctypes.data)
        return out

    def bin_results(self):
        arr = (PwppBinResult * self.nbins)()
        self._lib.simt_bin_results(self._h, C.byref(arr))
        return arr
