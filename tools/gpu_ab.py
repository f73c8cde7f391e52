"""A/B timing of the kernel-variant switches in ONE process (the switches are read when a context is created).
usage: python tools/gpu_ab.py [frames] [steps] -- prints one JSON line per configuration, also to gpurun_out/ab.jsonl.
Every configuration is also checked against the first one: identical ground/non-ground index lists for sampled frames."""
import hashlib, json, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "patchwork-plusplus_b200"))
import pwpp_b200, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
SWITCHES = ["PWPP_HIST_PIPE", "PWPP_SCATTER_V", "PWPP_SERIAL_FIT", "PWPP_S_MINB", "PWPP_M_MINB", "PWPP_L1_MINB", "PWPP_L2_MINB", "PWPP_L2_NW", "PWPP_L3_NW"]
CONFIGS = [
    {},
    {"PWPP_L3_NW": "16"},
    {"PWPP_L2_NW": "16"},
    {"PWPP_L3_NW": "16", "PWPP_L2_NW": "16"},
    {"PWPP_SCATTER_V": "1"},
    {},
]
pts, offs = synth.make_batch(20260922, 0, F, "kitti64", "cuda")
offs_np = offs.numpy()
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream; assert st != 0
ref_sig = None
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
out = open(os.path.join(REPO, "gpurun_out", "ab.jsonl"), "w")
for cfg in CONFIGS:
    for k in SWITCHES: os.environ.pop(k, None)
    os.environ.update(cfg)
    eng = pwpp_b200.Engine(device=0, num_streams=F)
    def step():
        eng.reset(); eng.estimate_device(pts.data_ptr(), offs_np, True, st)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    sig = hashlib.sha1()
    for f in (0, 1, F // 2, F - 1):
        sig.update(np.ascontiguousarray(eng.ground_indices(f)).tobytes()); sig.update(np.ascontiguousarray(eng.nonground_indices(f)).tobytes())
    tot_g = sum(eng.num_ground(f) for f in range(0, F, max(1, F // 64)))
    sig = sig.hexdigest()[:12]
    if ref_sig is None: ref_sig = sig
    eng.set_profiling(True)
    acc = None
    for _ in range(3):
        step(); eng.synchronize()
        t = eng.stage_times_ms()
        acc = t if acc is None else {k: acc[k] + v for k, v in t.items()}
    eng.set_profiling(False)
    stage = {k[2:]: round(v / 3, 3) for k, v in acc.items()}
    line = {"cfg": cfg, "ms_per_step": round(ms, 4), "fps": round(F / ms * 1e3, 1), "same_as_first": sig == ref_sig, "sig": sig, "ground_sampled": tot_g, "stage_ms": stage}
    print(json.dumps(line), flush=True); out.write(json.dumps(line) + "\n"); out.flush()
    eng.close(); del eng
torch.cuda.synchronize()
out.close()
os._exit(0)
