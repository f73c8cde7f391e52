"""Per-step CUDA-event times + wall clock + SM clock over a long run (diagnostic)."""
import os, sys, time, subprocess
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "patchwork-plusplus_b200"))
import pwpp_b200, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
pts, offs = synth.make_batch(20260922, 0, F, "kitti64", "cuda")
offs_np = offs.numpy()
eng = pwpp_b200.Engine(device=0, num_streams=F)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream; assert st != 0
def step():
    eng.reset(); eng.estimate_device(pts.data_ptr(), offs_np, True, st)
for _ in range(3): step()
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
t0 = time.perf_counter()
evs[0].record()
for k in range(K):
    step(); evs[k + 1].record()
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
per = [evs[k].elapsed_time(evs[k + 1]) for k in range(K)]
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print(f"F={F} K={K} events total {evs[0].elapsed_time(evs[K]):.2f} ms  wall {t_all*1e3:.2f} ms  host enqueue {t_enq*1e3:.2f} ms  clocks now: {clk}")
print("per-step ms:", " ".join(f"{p:.2f}" for p in per))
# the same with a synchronize after every step
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(K):
    step(); torch.cuda.synchronize()
print(f"wall with sync per step: {(time.perf_counter()-t0)*1e3/K:.3f} ms/step")
eng.close()
