#!/usr/bin/env python
"""The reference's python/examples/demo_sequential.py:18-42 call sequence on the B200 engine, minus the Open3D window:
one `pypatchworkpp.patchworkpp` instance, the scans of a directory in file-name order, every getter after each frame.

    python examples/demo_sequential.py /path/to/kitti/velodyne [--device 0]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "patchwork-plusplus_b200", "lib"))
import pypatchworkpp  # noqa: E402


def read_bin(bin_path):
    return np.fromfile(bin_path, dtype=np.float32).reshape((-1, 4))


if __name__ == "__main__":
    data_dir = sys.argv[1]
    device = int(sys.argv[sys.argv.index("--device") + 1]) if "--device" in sys.argv else 0
    params = pypatchworkpp.Parameters()
    params.verbose = False
    pw = pypatchworkpp.patchworkpp(params, device)
    for name in sorted(f for f in os.listdir(data_dir) if f.endswith(".bin")):
        cloud = read_bin(os.path.join(data_dir, name))
        pw.estimateGround(cloud)
        ground, nonground = pw.getGround(), pw.getNonground()
        gi, ngi = pw.getGroundIndices(), pw.getNongroundIndices()
        centers, normals = pw.getCenters(), pw.getNormals()
        print(f"{name}: points {cloud.shape[0]}  ground {ground.shape[0]}  nonground {nonground.shape[0]}  patches {centers.shape[0]}  "
              f"height {pw.getHeight():.4f}  time {pw.getTimeTaken() / 1e6:.6f} s")
