"""The C-ABI shared library and the drop-in surfaces load and export everything they declare (no compute calls)."""
import ctypes as C
import os
import re

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _declared_symbols():
    txt = open(os.path.join(REPO, "include", "pwpp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pwpp_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import pwpp_b200
    lib = pwpp_b200.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libpwpp_b200.so does not export {s}"
    assert lib.pwpp_abi_version() == 1


def test_params_default_matches_reference_defaults():
    import pwpp_b200
    from pwpp_ctypes import PwppParams, default_params
    lib = pwpp_b200.load_library()
    p = PwppParams()
    lib.pwpp_params_default(C.byref(p))
    q = default_params()
    for name, _ in PwppParams._fields_:
        a, b = getattr(p, name), getattr(q, name)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), name
        else:
            assert a == b, name


def test_create_fails_loudly_without_gpu_or_with_bad_params():
    import torch
    import pwpp_b200
    from pwpp_ctypes import default_params
    if not torch.cuda.is_available():
        with pytest.raises(pwpp_b200.PwppError, match="no CUDA device"):
            pwpp_b200.Engine()
    else:
        p = default_params(); p.num_zones = 5
        with pytest.raises(pwpp_b200.PwppError, match="num_zones"):
            pwpp_b200.Engine(p)


def test_python_module_surface_matches_reference_binding():
    """reference python/patchworkpp/pybinding.cpp:14-55"""
    import pypatchworkpp as m
    assert hasattr(m, "__version__")
    P = m.Parameters()
    for f in ["sensor_height", "verbose", "enable_RNR", "enable_RVPF", "enable_TGR", "num_iter", "num_lpr", "num_min_pts", "num_zones",
              "num_rings_of_interest", "RNR_ver_angle_thr", "RNR_intensity_thr", "th_seeds", "th_dist", "th_seeds_v", "th_dist_v",
              "max_range", "min_range", "uprightness_thr", "adaptive_seed_selection_margin", "intensity_thr", "num_sectors_each_zone",
              "num_rings_each_zone", "max_flatness_storage", "max_elevation_storage", "elevation_thr", "flatness_thr"]:
        assert hasattr(P, f), f
    assert (P.sensor_height, P.num_iter, P.num_lpr, P.num_min_pts, P.th_dist_v) == (1.723, 3, 20, 10, 0.1)
    assert P.num_sectors_each_zone == [16, 32, 54, 32] and P.num_rings_each_zone == [2, 4, 4, 4]
    for meth in ["getHeight", "getTimeTaken", "getGround", "getNonground", "getCenters", "getGroundIndices", "getNongroundIndices",
                 "getNormals", "estimateGround"]:
        assert hasattr(m.patchworkpp, meth), meth


def test_cpp_header_compiles_with_and_without_eigen(tmp_path):
    """include/patchwork/patchworkpp.h must compile standalone, and with an Eigen on the include path expose the
    reference's Eigen signatures (checked against oracle/eigen_shim, the only Eigen-like headers in the container)."""
    import subprocess
    src = tmp_path / "t.cpp"
    src.write_text('#include "patchwork/patchworkpp.h"\n'
                   'int main(){ patchwork::Params p; return p.num_zones == 4 ? 0 : 1; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(REPO, "include"), str(src)])
    src2 = tmp_path / "t2.cpp"
    src2.write_text('#include "patchwork/patchworkpp.h"\n'
                    '#ifndef PATCHWORKPP_HAVE_EIGEN\n#error Eigen branch not taken\n#endif\n'
                    'void f(patchwork::PatchWorkpp& pw, Eigen::MatrixXf c){ pw.estimateGround(c); Eigen::MatrixX3f g = pw.getGround();'
                    ' Eigen::VectorXi i = pw.getGroundIndices(); Eigen::MatrixX3f n = pw.getNormals(); (void)g;(void)i;(void)n;'
                    ' double h = pw.getHeight() + pw.getTimeTaken(); (void)h; }\nint main(){return 0;}\n')
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(REPO, "include"),
                           "-I" + os.path.join(REPO, "oracle", "eigen_shim"), str(src2)])
