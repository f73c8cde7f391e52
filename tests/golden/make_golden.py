"""Generates the committed golden fixtures. Runs ONLY in the build container (needs /root/reference).

Inputs : the reference's six KITTI scans, /root/reference/data/00000{0..5}.bin (N x 4 float32).
Outputs: tests/golden/kitti_00000X.npz  the scans themselves (column-major for better compression; the GPU
                                        box has no /root/reference)
         tests/golden/golden_ref.npz    outputs of the reference's OWN estimateGround — its patchworkpp.cpp
                                        compiled against oracle/eigen_shim (oracle/_ref/libpwref_stable.so: per-bin sort
                                        made stable so that fp32 sums are reproducible, see oracle/ref_capi.cpp; the
                                        unmodified build libpwref.so is asserted to give the same label sets) — for three parameter sets x {fresh instance per
                                        scan, one instance over the six scans in order}:
                                          ground masks (bit-packed), emission-ordered index lists' hashes,
                                          centers, normals, adaptive state after every scan.
Usage: python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, os.path.join(REPO, "tests"))
import oracle_py as O  # noqa: E402
from param_sets import PARAM_SETS  # noqa: E402

REF_DATA = "/root/reference/data"


def main():
    O.build()
    scans = []
    for f in range(6):
        a = np.fromfile(os.path.join(REF_DATA, f"{f:06d}.bin"), dtype=np.float32).reshape(-1, 4)
        scans.append(a)
        np.savez_compressed(os.path.join(HERE, f"kitti_{f:06d}.npz"), xyzi_t=np.ascontiguousarray(a.T))
    out = {}
    for pname, (mk, cols) in PARAM_SETS.items():
        for mode in ("fresh", "seq"):
            ref = O.Reference(mk(), stable_sort=True)
            raw = O.Reference(mk(), stable_sort=False)  # the unmodified reference (std::sort)
            for f, a in enumerate(scans):
                if mode == "fresh":
                    ref = O.Reference(mk(), stable_sort=True)
                    raw = O.Reference(mk(), stable_sort=False)
                ref.estimate(a[:, :cols])
                raw.estimate(a[:, :cols])
                # tie order of the per-bin sort must not change any label
                assert np.array_equal(np.sort(raw.getGroundIndices()), np.sort(ref.getGroundIndices()))
                g = ref.getGroundIndices()
                ng = ref.getNongroundIndices()
                mask = np.zeros(a.shape[0], dtype=bool)
                mask[g] = True
                assert len(g) + len(ng) == a.shape[0] and len(np.unique(np.concatenate([g, ng]))) == a.shape[0]
                st = ref.state()
                k = f"{pname}/{mode}/{f}"
                out[k + "/ground_mask"] = np.packbits(mask)
                out[k + "/n_ground"] = np.int64(len(g))
                out[k + "/centers"] = ref.getCenters()
                out[k + "/normals"] = ref.getNormals()
                out[k + "/state"] = np.array([st.sensor_height, *st.elevation_thr, *st.flatness_thr], dtype=np.float64)
                out[k + "/hist_n"] = np.array([*st.n_elevation, *st.n_flatness], dtype=np.int32)
                print(k, len(g), len(ng), ref._f("num_patches")(ref._h), st.sensor_height)
    np.savez_compressed(os.path.join(HERE, "golden_ref.npz"), **out)


if __name__ == "__main__":
    main()
