"""Thin ctypes front-end of the C-ABI in include/pwpp.h (lib/libpwpp_b200.so).

Used by tests/ and bench.py to call the product exactly as a foreign-language host would: plain
pointers and sizes, no torch types. The CUDA library is REQUIRED: loading fails loudly when it is
missing and creating an Engine fails loudly without a CUDA device — there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from pwpp_ctypes import PwppBinResult, PwppParams, PwppState, default_params  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PWPP_LIB") or os.path.join(_HERE, "lib", "libpwpp_b200.so")   # PWPP_LIB: a diagnostic build of the same library

_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python patchwork-plusplus_b200/build.py` "
                           "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.pwpp_params_default.argtypes = [C.POINTER(PwppParams)]; lib.pwpp_params_default.restype = None
    lib.pwpp_create.argtypes = [C.POINTER(PwppParams), i32, i32, i64, C.POINTER(vp)]; lib.pwpp_create.restype = i32
    lib.pwpp_destroy.argtypes = [vp]; lib.pwpp_destroy.restype = None
    lib.pwpp_last_error.argtypes = []; lib.pwpp_last_error.restype = C.c_char_p
    lib.pwpp_abi_version.argtypes = []; lib.pwpp_abi_version.restype = i32
    lib.pwpp_num_bins.argtypes = [vp]; lib.pwpp_num_bins.restype = i32
    lib.pwpp_estimate_host.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i64), i32, i64, i64]; lib.pwpp_estimate_host.restype = i32
    lib.pwpp_estimate_device.argtypes = [vp, i32, vp, C.POINTER(i64), i32, vp]; lib.pwpp_estimate_device.restype = i32
    lib.pwpp_synchronize.argtypes = [vp]; lib.pwpp_synchronize.restype = i32
    for n in ("pwpp_num_ground", "pwpp_num_nonground"):
        getattr(lib, n).argtypes = [vp, i32]; getattr(lib, n).restype = i64
    for n in ("pwpp_copy_ground_indices", "pwpp_copy_nonground_indices", "pwpp_copy_ground_xyz", "pwpp_copy_nonground_xyz",
              "pwpp_copy_centers", "pwpp_copy_normals", "pwpp_copy_bin_results", "pwpp_copy_bin_ids"):
        getattr(lib, n).argtypes = [vp, i32, vp]; getattr(lib, n).restype = i32
    lib.pwpp_num_patches.argtypes = [vp, i32]; lib.pwpp_num_patches.restype = i32
    lib.pwpp_height.argtypes = [vp, i32]; lib.pwpp_height.restype = C.c_double
    lib.pwpp_time_us.argtypes = [vp]; lib.pwpp_time_us.restype = C.c_double
    lib.pwpp_call_times_us.argtypes = [vp, C.POINTER(C.c_float)]; lib.pwpp_call_times_us.restype = i32
    lib.pwpp_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]; lib.pwpp_device_results.restype = i32
    lib.pwpp_get_state.argtypes = [vp, i32, C.POINTER(PwppState)]; lib.pwpp_get_state.restype = i32
    lib.pwpp_copy_history.argtypes = [vp, i32, i32, i32, vp]; lib.pwpp_copy_history.restype = i32
    lib.pwpp_state_blob_size.argtypes = [vp]; lib.pwpp_state_blob_size.restype = C.c_size_t
    lib.pwpp_export_state.argtypes = [vp, i32, vp]; lib.pwpp_export_state.restype = i32
    lib.pwpp_import_state.argtypes = [vp, i32, vp, C.c_size_t]; lib.pwpp_import_state.restype = i32
    lib.pwpp_reset_stream.argtypes = [vp, i32]; lib.pwpp_reset_stream.restype = i32
    lib.pwpp_reset_all.argtypes = [vp]; lib.pwpp_reset_all.restype = i32
    lib.pwpp_host_alloc.argtypes = [C.c_size_t]; lib.pwpp_host_alloc.restype = vp
    lib.pwpp_host_free.argtypes = [vp]; lib.pwpp_host_free.restype = None
    lib.pwpp_set_profiling.argtypes = [vp, i32]; lib.pwpp_set_profiling.restype = i32
    lib.pwpp_stage_times_ms.argtypes = [vp, C.POINTER(C.c_float)]; lib.pwpp_stage_times_ms.restype = i32
    lib.pwpp_stage_name.argtypes = [i32]; lib.pwpp_stage_name.restype = C.c_char_p
    lib.pwpp_launch_count.argtypes = [vp]; lib.pwpp_launch_count.restype = i64
    lib.pwpp_set_output_order.argtypes = [vp, i32]; lib.pwpp_set_output_order.restype = i32
    lib.pwpp_host_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]; lib.pwpp_host_results.restype = i32
    lib.pwpp_bind_host_to_device.argtypes = [i32]; lib.pwpp_bind_host_to_device.restype = i32
    _lib = lib
    return lib


def bind_host_to_device(device: int) -> int:
    """pwpp_bind_host_to_device: pin the calling thread to the CPUs of the GPU's NUMA node (returns the node or -1)."""
    return int(load_library().pwpp_bind_host_to_device(device))


class PwppError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise PwppError(f"pwpp error {rc}: {load_library().pwpp_last_error().decode()}")


class Engine:
    """One `pwpp_ctx`: `num_streams` independent sensor streams on one CUDA device."""

    def __init__(self, params: PwppParams = None, device: int = 0, num_streams: int = 1, max_points_per_frame: int = 0):
        self.lib = load_library()
        self.params = params if params is not None else default_params()
        h = C.c_void_p()
        _check(self.lib.pwpp_create(C.byref(self.params), device, num_streams, max_points_per_frame, C.byref(h)))
        self._h = h
        self.num_streams = num_streams
        self.nbins = self.lib.pwpp_num_bins(h)
        self._n = []

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pwpp_destroy(self._h)
            self._h = None

    def __del__(self):
        # at interpreter shutdown the CUDA runtime may already be torn down: only release explicitly-open handles
        # while the library is still importable
        try:
            import sys
            if sys is not None and not sys.is_finalizing():
                self.close()
        except Exception:
            pass

    # ---- hot path ----
    def estimate_host(self, frames):
        """frames: list of C-contiguous float32 arrays (n_f, 3|4), one per stream."""
        frames = [np.ascontiguousarray(f, dtype=np.float32) for f in frames]
        cols = frames[0].shape[1]
        assert all(f.ndim == 2 and f.shape[1] == cols for f in frames)
        nf = len(frames)
        ptrs = (C.c_void_p * nf)(*[f.ctypes.data for f in frames])
        ns = (C.c_int64 * nf)(*[f.shape[0] for f in frames])
        self._n = [f.shape[0] for f in frames]
        _check(self.lib.pwpp_estimate_host(self._h, nf, ptrs, ns, cols, cols, 1))

    def estimate_host_strided(self, ptrs, ns, cols, row_stride, col_stride):
        nf = len(ptrs)
        p = (C.c_void_p * nf)(*ptrs)
        n = (C.c_int64 * nf)(*ns)
        self._n = list(ns)
        _check(self.lib.pwpp_estimate_host(self._h, nf, p, n, cols, row_stride, col_stride))

    def estimate_device(self, d_ptr: int, offsets, has_intensity: bool = True, stream: int = 0):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        nf = len(offsets) - 1
        self._n = np.diff(offsets).tolist()
        _check(self.lib.pwpp_estimate_device(self._h, nf, C.c_void_p(d_ptr), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                             1 if has_intensity else 0, C.c_void_p(stream)))

    def synchronize(self):
        _check(self.lib.pwpp_synchronize(self._h))

    # ---- results ----
    def num_ground(self, f=0): return int(self.lib.pwpp_num_ground(self._h, f))
    def num_nonground(self, f=0): return int(self.lib.pwpp_num_nonground(self._h, f))

    def _get(self, fn, f, n, dtype, shape):
        out = np.empty(shape, dtype=dtype)
        if n > 0:
            _check(fn(self._h, f, out.ctypes.data))
        return out

    def ground_indices(self, f=0):
        n = self.num_ground(f); return self._get(self.lib.pwpp_copy_ground_indices, f, n, np.int32, (n,))

    def nonground_indices(self, f=0):
        n = self.num_nonground(f); return self._get(self.lib.pwpp_copy_nonground_indices, f, n, np.int32, (n,))

    def ground_xyz(self, f=0):
        n = self.num_ground(f); return self._get(self.lib.pwpp_copy_ground_xyz, f, n, np.float32, (n, 3))

    def nonground_xyz(self, f=0):
        n = self.num_nonground(f); return self._get(self.lib.pwpp_copy_nonground_xyz, f, n, np.float32, (n, 3))

    def num_patches(self, f=0): return int(self.lib.pwpp_num_patches(self._h, f))

    def centers(self, f=0):
        n = self.num_patches(f); return self._get(self.lib.pwpp_copy_centers, f, n, np.float32, (n, 3))

    def normals(self, f=0):
        n = self.num_patches(f); return self._get(self.lib.pwpp_copy_normals, f, n, np.float32, (n, 3))

    def height(self, f=0): return float(self.lib.pwpp_height(self._h, f))
    def time_us(self): return float(self.lib.pwpp_time_us(self._h))

    def call_times_us(self):
        """{h2d, kernels, d2h, device_total} of the last single-chunk estimate_host call, microseconds (CUDA events)."""
        arr = (C.c_float * 4)()
        _check(self.lib.pwpp_call_times_us(self._h, arr))
        return dict(zip(("h2d", "kernels", "d2h", "device_total"), (float(v) for v in arr)))

    def bin_results(self, f=0):
        arr = (PwppBinResult * self.nbins)()
        _check(self.lib.pwpp_copy_bin_results(self._h, f, C.byref(arr)))
        return arr

    def bin_ids(self, f=0):
        out = np.empty(self._n[f], dtype=np.uint16)
        if self._n[f] > 0:
            _check(self.lib.pwpp_copy_bin_ids(self._h, f, out.ctypes.data))
        return out

    def state(self, f=0) -> PwppState:
        st = PwppState()
        _check(self.lib.pwpp_get_state(self._h, f, C.byref(st)))
        return st

    def history(self, f, ring, which):
        st = self.state(f)
        n = (st.n_flatness if which else st.n_elevation)[ring]
        out = np.empty(n, dtype=np.float64)
        _check(self.lib.pwpp_copy_history(self._h, f, ring, which, out.ctypes.data))
        return out

    def export_state(self, f=0) -> bytes:
        """Complete temporal state of stream f as an opaque blob (checkpoint / migration to another ctx or GPU)."""
        buf = C.create_string_buffer(self.lib.pwpp_state_blob_size(self._h))
        _check(self.lib.pwpp_export_state(self._h, f, buf))
        return buf.raw

    def import_state(self, f, blob: bytes):
        _check(self.lib.pwpp_import_state(self._h, f, blob, len(blob)))

    def device_index_lists(self):
        """Zero-copy view of the last call's results on the device: (indices, num_ground) as torch int32 CUDA tensors.
        indices is laid out like the input (frame f's region starts at its point offset): ground list, then non-ground
        list. Valid until the next estimate call; the caller synchronizes with its stream (or Engine.synchronize())."""
        import torch
        d_idx, d_ng = self.device_results()
        total, nf = int(sum(self._n)), len(self._n)

        class _View:   # __cuda_array_interface__ v3: torch.as_tensor wraps the memory without copying
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3, "strides": None}
        idx = torch.as_tensor(_View(d_idx, total), device="cuda") if total > 0 else torch.empty(0, dtype=torch.int32, device="cuda")
        ng = torch.as_tensor(_View(d_ng, nf), device="cuda")
        return idx, ng

    def host_index_lists(self):
        """Zero-copy numpy views of the last call's results in the page-locked result buffer: (indices int32[total],
        num_ground int32[nframes], offsets int64[nframes + 1]). Frame f: ground = indices[off[f] : off[f] + ng[f]], non-ground
        follows up to off[f + 1] (minus dropped points). Valid until the next estimate call."""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(self.lib.pwpp_host_results(self._h, C.byref(a), C.byref(b), C.byref(c)))
        total, nf = int(sum(self._n)), len(self._n)
        idx = np.ctypeslib.as_array((C.c_int32 * max(total, 1)).from_address(a.value))[:total]
        ng = np.ctypeslib.as_array((C.c_int32 * nf).from_address(b.value))
        off = np.ctypeslib.as_array((C.c_int64 * (nf + 1)).from_address(c.value))
        return idx, ng, off

    def device_results(self):
        a, b = C.c_void_p(), C.c_void_p()
        _check(self.lib.pwpp_device_results(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    ORDER_BIN, ORDER_REFERENCE = 0, 1

    def set_output_order(self, order: int):
        """ORDER_BIN (default): ascending point index inside a bin; ORDER_REFERENCE: the reference's z order (pwpp.h)."""
        _check(self.lib.pwpp_set_output_order(self._h, order))

    NUM_STAGES = 11

    def set_profiling(self, on: bool):
        _check(self.lib.pwpp_set_profiling(self._h, 1 if on else 0))

    def stage_times_ms(self):
        arr = (C.c_float * self.NUM_STAGES)()
        _check(self.lib.pwpp_stage_times_ms(self._h, arr))
        return {self.lib.pwpp_stage_name(i).decode(): float(arr[i]) for i in range(self.NUM_STAGES)}

    def launch_count(self) -> int:
        return int(self.lib.pwpp_launch_count(self._h))

    def reset(self, f=None):
        _check(self.lib.pwpp_reset_all(self._h) if f is None else self.lib.pwpp_reset_stream(self._h, f))
