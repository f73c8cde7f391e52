"""Diagnostic: one batch of synthetic frames with the reference emission order on (k_order in the launch sequence); used under ncu
(-k regex:k_order) and stand-alone (prints the step time with and without k_order)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "patchwork-plusplus_b200")]
import numpy as np, torch
import pwpp_b200, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
pts, offs = synth.make_batch(20260922, 0, F, "kitti64", dev)
offs_np = offs.numpy()
eng = pwpp_b200.Engine(device=0, num_streams=F, max_points_per_frame=int(np.diff(offs_np).max()))
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
for order in (0, 1):
    eng.set_output_order(order)
    for _ in range(3):
        eng.reset(); eng.estimate_device(pts.data_ptr(), offs_np, True, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.reset(); eng.estimate_device(pts.data_ptr(), offs_np, True, st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    print("order", order, "frames", F, "ms/step", round(e0.elapsed_time(e1) / 5, 3))
eng.close()
