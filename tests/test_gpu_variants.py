"""Opt-in GPU check of the switched-off kernel variants (round-2 candidates): every variant must return exactly what the
default configuration returns. NOT part of the default `-m gpu` run: these variants have only been executed on the SIMT
twin so far (tests/test_simt_kernels.py); run it explicitly, under a timeout, with

    PWPP_TEST_VARIANTS=1 timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -q

tools/gpu_round_end.sh runs every variant in its OWN process under its own timeout (a variant that hangs the GPU kills only
its process) and hands the list of variants that passed to tools/gpu_tune.py.
"""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("PWPP_TEST_VARIANTS") != "1", reason="opt-in: set PWPP_TEST_VARIANTS=1")]

VARIANTS = [
    {"PWPP_FRONT": "1"},
    {"PWPP_PART_ILP": "1"},
    {"PWPP_EMIT_SPLIT": "8"},
    {"PWPP_SOLVE_CALL": "1"},
    {"PWPP_FUSE_SEED": "0", "PWPP_L2_MINB": "4"},
    {"PWPP_FUSE_SEED": "3"},
    {"PWPP_L2_WIDE": "1"},
    {"PWPP_X_FIXPOINT": "1"},
    {"PWPP_M_RESIDENT": "1"},
    {"PWPP_L1_CTA": "1"},
    {"PWPP_M_HALF": "1"},
    {"PWPP_L2_PLS": "1"},
    {"PWPP_L2_PLS": "1", "PWPP_L2_MINB": "4"},
    {"PWPP_FRONT": "1", "PWPP_PART_ILP": "1", "PWPP_EMIT_SPLIT": "4", "PWPP_SOLVE_CALL": "1", "PWPP_L2_WIDE": "1"},
]
SWITCHES = sorted({k for v in VARIANTS for k in v})


def _frames(kitti):
    import synth
    rng = np.random.default_rng(5)
    big = np.c_[5 + rng.random(20000) * 0.5, rng.random(20000) * 0.5, -1.7 + rng.normal(0, 0.02, 20000), rng.random(20000)].astype(np.float32)
    return [kitti[0], kitti[3], np.zeros((0, 4), np.float32), synth.make_frame(9, 0).numpy(), big, kitti[5][:4097], synth.make_frame(5, 0, "dense1m", "cuda").cpu().numpy()]


def _run(env, frames):
    import pwpp_b200
    saved = {k: os.environ.pop(k, None) for k in SWITCHES}
    os.environ.update(env)
    try:
        eng = pwpp_b200.Engine(device=0, num_streams=len(frames))   # the switches are read here
        out = []
        for rep in range(2):                                        # two consecutive frames per stream: temporal state too
            eng.estimate_host(frames)
            out.append([(eng.ground_indices(f).copy(), eng.nonground_indices(f).copy(), bytes(eng.bin_results(f)), bytes(eng.state(f))) for f in range(len(frames))])
        eng.close()
        return out
    finally:
        for k in SWITCHES:
            os.environ.pop(k, None)
            if saved[k] is not None:
                os.environ[k] = saved[k]


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join(f"{k[5:]}={v}" for k, v in e.items()))
def test_variant_equals_default(kitti, env):
    frames = _frames(kitti)
    base, var = _run({}, frames), _run(env, frames)
    for rep in range(2):
        for f in range(len(frames)):
            assert np.array_equal(base[rep][f][0], var[rep][f][0]) and np.array_equal(base[rep][f][1], var[rep][f][1]), f"{env}: index lists differ, call {rep} frame {f}"
            rounding_only = any(k in env for k in ("PWPP_X_FIXPOINT", "PWPP_M_RESIDENT", "PWPP_L1_CTA", "PWPP_M_HALF"))
            if env.get("PWPP_FUSE_SEED") != "0" and not rounding_only:
                assert base[rep][f][2] == var[rep][f][2], f"{env}: patch records differ, call {rep} frame {f}"
            if not rounding_only:
                assert base[rep][f][3] == var[rep][f][3], f"{env}: stream state differs, call {rep} frame {f}"
