"""Parameter sets exercised by the parity tests.

  default : the reference's Params() defaults (patchworkpp.h:79-111), N x 4 input
  ros     : the overrides of the reference's ROS2 launch file (ros/launch/patchworkpp.launch.py:50-64 and
            ros/src/GroundSegmentationServer.cpp:47: enable_RNR = false), N x 3 input like
            ros/src/Utils.hpp:158-172 produces
  no_rvpf_tgr : defaults with R-VPF and TGR switched off and a different bin layout
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "patchwork-plusplus_b200"))
from pwpp_ctypes import default_params  # noqa: E402


def _default():
    return default_params()


def _ros():
    p = default_params()
    p.sensor_height = 1.88
    p.num_iter = 3
    p.num_lpr = 20
    p.num_min_pts = 0
    p.th_seeds = 0.3
    p.th_dist = 0.125
    p.th_seeds_v = 0.25
    p.th_dist_v = 0.9
    p.max_range = 80.0
    p.min_range = 1.0
    p.uprightness_thr = 0.101
    p.enable_RNR = 0
    return p


def _no_rvpf_tgr():
    p = default_params()
    p.enable_RVPF = 0
    p.enable_TGR = 0
    p.num_sectors_each_zone[:] = [12, 24, 40, 20]
    p.num_rings_each_zone[:] = [3, 3, 5, 2]
    p.num_lpr = 12
    p.num_min_pts = 6
    p.num_rings_of_interest = 3
    p.max_flatness_storage = 40
    p.max_elevation_storage = 50
    return p


PARAM_SETS = {
    "default": (_default, 4),
    "ros": (_ros, 3),
    "no_rvpf_tgr": (_no_rvpf_tgr, 4),
}
