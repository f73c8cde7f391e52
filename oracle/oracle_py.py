"""ctypes front-ends of the CPU checkers (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.

  Oracle(params, arith)  -> oracle/_ref/libpwpp_oracle.so   (plain-C restatement, pwpp_oracle.c)
  Reference(params, stable_sort) -> oracle/_ref/libpwref[_stable].so (the reference's own
                            patchworkpp.cpp compiled against oracle/eigen_shim, ref_capi.cpp)

Both expose the same tiny interface: estimate(points) then the getters of the reference class
(reference python/patchworkpp/pybinding.cpp:45-55) plus state()/history() for the temporal state.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
sys.path.insert(0, os.path.join(_REPO, "patchwork-plusplus_b200"))
from pwpp_ctypes import PwppParams, PwppState, PwppBinResult, default_params  # noqa: E402

ARITH_REF32 = 0
ARITH_CANON64 = 1


def build(force: bool = False) -> None:
    """Build the checkers (gcc/g++ only). The reference-derived libs are rebuilt only where
    /root/reference exists; on the GPU box the prebuilt oracle/_ref/*.so travel with the snapshot."""
    out = os.path.join(_HERE, "_ref")
    need = force or not os.path.exists(os.path.join(out, "libpwpp_oracle.so")) \
        or os.path.getmtime(os.path.join(out, "libpwpp_oracle.so")) < os.path.getmtime(os.path.join(_HERE, "pwpp_oracle.c"))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    have_ref = os.path.exists(os.path.join(out, "libpwref.so")) and os.path.exists(os.path.join(out, "libpwref_stable.so"))
    if os.path.exists("/root/reference/cpp/patchworkpp/src/patchworkpp.cpp") and (force or not have_ref):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _load(name):
    path = os.path.join(_HERE, "_ref", name)
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


class _Base:
    _prefix = None
    _lib = None

    def _bind(self, lib, pfx):
        vp = C.c_void_p
        f = lambda n: getattr(lib, pfx + n)  # noqa: E731
        f("estimate").argtypes = [vp, vp, C.c_int64, C.c_int]
        f("estimate").restype = None
        for n in ("num_ground", "num_nonground"):
            f(n).argtypes = [vp]; f(n).restype = C.c_int64
        for n in ("ground_indices", "nonground_indices", "ground_xyz", "nonground_xyz", "centers", "normals"):
            if hasattr(lib, pfx + n):
                f(n).argtypes = [vp, vp]; f(n).restype = None
        f("num_patches").argtypes = [vp]; f("num_patches").restype = C.c_int
        if hasattr(lib, pfx + "height"):
            f("height").argtypes = [vp]; f("height").restype = C.c_double
        f("get_state").argtypes = [vp, C.POINTER(PwppState)]; f("get_state").restype = None
        f("history").argtypes = [vp, C.c_int, C.c_int, vp]; f("history").restype = None
        f("destroy").argtypes = [vp]; f("destroy").restype = None
        self._f = f

    def estimate(self, pts: np.ndarray) -> None:
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        assert pts.ndim == 2 and pts.shape[1] in (3, 4)
        self._n = pts.shape[0]
        self._f("estimate")(self._h, pts.ctypes.data, pts.shape[0], pts.shape[1])

    def _ivec(self, name, n):
        out = np.empty(n, dtype=np.int32)
        self._f(name)(self._h, out.ctypes.data)
        return out

    def _x3(self, name, n):
        out = np.empty((n, 3), dtype=np.float32)
        self._f(name)(self._h, out.ctypes.data)
        return out

    def getGroundIndices(self): return self._ivec("ground_indices", self._f("num_ground")(self._h))
    def getNongroundIndices(self): return self._ivec("nonground_indices", self._f("num_nonground")(self._h))
    def getGround(self): return self._x3("ground_xyz", self._f("num_ground")(self._h))
    def getNonground(self): return self._x3("nonground_xyz", self._f("num_nonground")(self._h))
    def getCenters(self): return self._x3("centers", self._f("num_patches")(self._h))
    def getNormals(self): return self._x3("normals", self._f("num_patches")(self._h))
    def getHeight(self): return self._f("height")(self._h)

    def state(self) -> PwppState:
        st = PwppState()
        self._f("get_state")(self._h, C.byref(st))
        return st

    def history(self, ring: int, which: int) -> np.ndarray:
        st = self.state()
        n = (st.n_flatness if which else st.n_elevation)[ring]
        out = np.empty(n, dtype=np.float64)
        self._f("history")(self._h, ring, which, out.ctypes.data)
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._f("destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Oracle(_Base):
    """oracle/pwpp_oracle.c. arith: ARITH_REF32 (literal reference precisions) or ARITH_CANON64."""

    def __init__(self, params: PwppParams = None, arith: int = ARITH_CANON64):
        lib = _load("libpwpp_oracle.so")
        self._bind(lib, "pwo_")
        lib.pwo_create.argtypes = [C.POINTER(PwppParams), C.c_int]; lib.pwo_create.restype = C.c_void_p
        lib.pwo_num_bins.argtypes = [C.c_void_p]; lib.pwo_num_bins.restype = C.c_int
        lib.pwo_bin_ids.argtypes = [C.c_void_p, C.c_void_p]; lib.pwo_bin_ids.restype = None
        lib.pwo_bin_results.argtypes = [C.c_void_p, C.c_void_p]; lib.pwo_bin_results.restype = None
        lib.pwo_bin_min_fit_n.argtypes = [C.c_void_p, C.c_void_p]; lib.pwo_bin_min_fit_n.restype = None
        self._lib = lib
        self.params = params if params is not None else default_params()
        self._h = lib.pwo_create(C.byref(self.params), arith)
        if not self._h:
            raise ValueError("oracle: unsupported parameters (num_zones != 4 or num_rings_of_interest > 4)")
        self.nbins = lib.pwo_num_bins(self._h)

    def bin_ids(self) -> np.ndarray:
        out = np.empty(self._n, dtype=np.uint16)
        self._lib.pwo_bin_ids(self._h, out.ctypes.data)
        return out

    def bin_min_fit_n(self) -> np.ndarray:
        """Per bin: smallest non-empty point set given to estimate_plane (INT32_MAX if never fitted)."""
        out = np.empty(self.nbins, dtype=np.int32)
        self._lib.pwo_bin_min_fit_n(self._h, out.ctypes.data)
        return out

    def bin_results(self):
        arr = (PwppBinResult * self.nbins)()
        self._lib.pwo_bin_results(self._h, C.byref(arr))
        return arr


class Reference(_Base):
    """The reference's own estimateGround (patchworkpp.cpp compiled against oracle/eigen_shim)."""

    def __init__(self, params: PwppParams = None, stable_sort: bool = False):
        lib = _load("libpwref_stable.so" if stable_sort else "libpwref.so")
        self._bind(lib, "pwref_")
        lib.pwref_create.argtypes = [C.POINTER(PwppParams)]; lib.pwref_create.restype = C.c_void_p
        lib.pwref_time_taken.argtypes = [C.c_void_p]; lib.pwref_time_taken.restype = C.c_double
        self._lib = lib
        self.params = params if params is not None else default_params()
        self._h = lib.pwref_create(C.byref(self.params))

    def getTimeTaken(self): return self._lib.pwref_time_taken(self._h)


def have_reference_build() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libpwref.so")) and \
        os.path.exists(os.path.join(_HERE, "_ref", "libpwref_stable.so"))
