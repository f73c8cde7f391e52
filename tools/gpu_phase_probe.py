"""Diagnostic: where the k_fit_group CTAs spend their cycles (needs a library built with -DPWPP_PHASE_CLOCKS, PWPP_LIB=...)."""
import ctypes as C, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "patchwork-plusplus_b200")]
import numpy as np, torch
import pwpp_b200, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pts, offs = synth.make_batch(20260922, 0, F, "kitti64", torch.device("cuda", 0))
eng = pwpp_b200.Engine(device=0, num_streams=F)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
lib = eng.lib
lib.pwpp_debug_phase_clocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
for _ in range(3):
    eng.reset(); eng.estimate_device(pts.data_ptr(), offs.numpy(), True, st.cuda_stream)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 48)()
lib.pwpp_debug_phase_clocks(eng.h if hasattr(eng, "h") else eng._h, buf, 1)
eng.reset(); eng.estimate_device(pts.data_ptr(), offs.numpy(), True, st.cuda_stream)
torch.cuda.synchronize()
lib.pwpp_debug_phase_clocks(eng.h if hasattr(eng, "h") else eng._h, buf, 0)
names = ["stage+setup", "selection", "pass", "solve", "-", "partition", "-", "-", "groups", "rounds", "seed_rounds"]
for c, cn in enumerate("ABC"):
    v = [buf[c * 16 + i] for i in range(12)]
    tot = sum(v[:6]) or 1
    g = max(v[8], 1)
    print(cn, {names[i]: f"{100 * v[i] / tot:.1f}% ({v[i] / g / 1965:.2f} us/group)" for i in (0, 1, 2, 3, 5)}, "groups", v[8], "rounds/group %.2f" % (v[9] / g), "seed rounds/group %.2f" % (v[10] / g), "us/group %.1f" % (tot / g / 1965))

# event trace of CTA 0 (class PWPP_TRACE_CLS of the build): per-warp timeline of the first groups
lib.pwpp_debug_events.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
ev = (C.c_uint * (4 * 16384))()
lib.pwpp_debug_events(eng._h, ev, 0)   # clear
eng.reset(); eng.estimate_device(pts.data_ptr(), offs.numpy(), True, st.cuda_stream)
torch.cuda.synchronize()
lib.pwpp_debug_events(eng._h, ev, 16384)
E = np.frombuffer(ev, dtype=np.uint32).reshape(-1, 4)
E = E[E[:, 3] == 1]
np.save(os.path.join(REPO, "gpurun_out", "events.npy"), E)
print("events", len(E))
