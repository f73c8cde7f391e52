"""Back-to-back (un-synchronised) estimate_device calls: results must still match the oracle, and CUDA-event timing
must agree with wall-clock timing around explicit synchronisation."""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "patchwork-plusplus_b200"), os.path.join(REPO, "oracle")):
    sys.path.insert(0, p)
import oracle_py as O, pwpp_b200, synth
F = 96
pts, offs = synth.make_batch(777, 0, F, "kitti64", "cuda")
pts2, offs2 = synth.make_batch(778, 0, F, "kitti64", "cuda")
offs_np, offs2_np = offs.numpy(), offs2.numpy()
eng = pwpp_b200.Engine(device=0, num_streams=F)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream; assert st != 0
# A, B, A back to back without any host synchronisation in between
for (p, o) in ((pts, offs_np), (pts2, offs2_np), (pts, offs_np)):
    eng.reset(); eng.estimate_device(p.data_ptr(), o, True, st)
eng.synchronize()
host = pts.cpu().numpy()
bad = 0
for f in range(0, F, 7):
    a = host[offs_np[f]:offs_np[f + 1]]
    orc = O.Oracle(arith=O.ARITH_CANON64); orc.estimate(a)
    if orc.bin_min_fit_n().min() < 3: continue
    g = np.sort(eng.ground_indices(f)); go = np.sort(orc.getGroundIndices())
    if not np.array_equal(g, go): bad += 1; print("MISMATCH frame", f, len(g), len(go))
print("back-to-back result check:", "OK" if bad == 0 else f"{bad} mismatches")
K = 20
def step():
    eng.reset(); eng.estimate_device(pts.data_ptr(), offs_np, True, st)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(K): step()
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"events: {e0.elapsed_time(e1)/K:.3f} ms/step   wall: {(t1-t0)*1e3/K:.3f} ms/step   ({F} frames/step)")
# per-step with a full sync after every step (no overlap between steps possible)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(K):
    step(); torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"wall with sync after every step: {(t1-t0)*1e3/K:.3f} ms/step")
eng.close()
