"""Deterministic synthetic LiDAR scans shaped like the reference's fixtures.

A frame is a ray-cast of a spinning multi-beam sensor at the origin against
  * a sloped ground plane  z = -h + sx*x + sy*y  (h ~ 1.73 m, |s| <= 0.03) with 2 cm range noise,
  * 20-40 axis-aligned boxes (cars 4x2x1.5, walls 0.3x20x3, poles 0.3x0.3x4) placed at r in [4, 60] m,
keeping returns up to 120 m, dropping 5 % at random, with intensity in {0, 0.01, ..., 0.99} and a handful of
low-intensity "reflection" points below the ground (exercise RNR, reference patchworkpp.cpp:377-400).

  kitti64   : 64 beams (+2 deg .. -24.33 deg, HDL-64E-like) x 2083 azimuth steps  -> ~110-125k points
  ouster128 : 128 beams uniform in +-22.5 deg x 8192 azimuth steps                -> ~0.5M returns (upward beams miss)
  dense1m   : the same beam layout x 16384 azimuth steps                          -> ~1.0M returns

torch is used only as the array engine (CPU here, CUDA on the GPU box where generating a 1024-frame batch
with numpy would take minutes). Frame f of seed s is fully determined by (s, f, sensor) on a given device.
"""
import math

import torch

SENSORS = {
    "kitti64": dict(n_az=2083),
    "ouster128": dict(n_az=8192),
    "dense1m": dict(n_az=16384),   # Ouster-128 beam layout, 2x azimuth density -> ~1.0M returns (BASELINE config 5)
}


def _elevations(sensor: str, device):
    if sensor == "kitti64":
        up = torch.linspace(2.0, -8.33, 32, device=device, dtype=torch.float64)
        lo = torch.linspace(-8.83, -24.33, 32, device=device, dtype=torch.float64)
        el = torch.cat([up, lo])
    elif sensor in ("ouster128", "dense1m"):
        el = torch.linspace(22.5, -22.5, 128, device=device, dtype=torch.float64)
    else:
        raise ValueError(sensor)
    return el * (math.pi / 180.0)


def make_frame(seed: int, f: int, sensor: str = "kitti64", device="cpu") -> torch.Tensor:
    """Returns one frame as a float32 tensor (n, 4) = x, y, z, intensity on `device`."""
    gen = torch.Generator(device=device)
    gen.manual_seed((seed * 1000003 + f * 7919 + 17) & 0x7FFFFFFF)
    R = lambda *shape: torch.rand(*shape, generator=gen, device=device, dtype=torch.float64)  # noqa: E731
    N = lambda *shape: torch.randn(*shape, generator=gen, device=device, dtype=torch.float64)  # noqa: E731
    n_az = SENSORS[sensor]["n_az"]
    el = _elevations(sensor, device)
    az = (torch.arange(n_az, device=device, dtype=torch.float64) + R(1)) * (2 * math.pi / n_az)
    # beam-major like a real spinning sensor's per-laser ordering
    E, A = torch.meshgrid(el, az, indexing="ij")
    E = E + N(*E.shape) * 2e-4
    dx, dy, dz = torch.cos(E) * torch.cos(A), torch.cos(E) * torch.sin(A), torch.sin(E)
    dx, dy, dz = dx.reshape(-1), dy.reshape(-1), dz.reshape(-1)
    h = 1.73 + (R(1) - 0.5) * 0.06
    sx, sy = (R(1) - 0.5) * 0.06, (R(1) - 0.5) * 0.06
    den = dz - sx * dx - sy * dy
    t_g = torch.where(den < -1e-6, -h / den, torch.full_like(den, float("inf")))
    # boxes
    nbox = int(20 + (R(1) * 21).item())
    kind = (R(nbox) * 3).floor()
    size = torch.where(kind[:, None] == 0, torch.tensor([4.0, 2.0, 1.5], device=device, dtype=torch.float64),
                       torch.where(kind[:, None] == 1, torch.tensor([0.3, 20.0, 3.0], device=device, dtype=torch.float64),
                                   torch.tensor([0.3, 0.3, 4.0], device=device, dtype=torch.float64)))
    swap = R(nbox) < 0.5
    size = torch.where(swap[:, None], size[:, [1, 0, 2]], size)
    br = 4.0 + R(nbox) * 56.0
    ba = R(nbox) * 2 * math.pi
    cx, cy = br * torch.cos(ba), br * torch.sin(ba)
    zb = -h + sx * cx + sy * cy
    lo = torch.stack([cx - size[:, 0] / 2, cy - size[:, 1] / 2, zb], dim=1)
    hi = torch.stack([cx + size[:, 0] / 2, cy + size[:, 1] / 2, zb + size[:, 2]], dim=1)
    d = torch.stack([dx, dy, dz], dim=1)  # (rays, 3)
    inv = 1.0 / torch.where(d.abs() < 1e-12, torch.full_like(d, 1e-12), d)
    t_best = t_g
    # loop over boxes keeps memory at O(rays) (1M-ray Ouster frames)
    for b in range(nbox):
        t0 = lo[b][None, :] * inv
        t1 = hi[b][None, :] * inv
        tmin = torch.minimum(t0, t1).amax(dim=1)
        tmax = torch.maximum(t0, t1).amin(dim=1)
        hit = (tmax >= tmin) & (tmax > 0)
        tb = torch.where(hit, torch.where(tmin > 0, tmin, tmax), torch.full_like(tmin, float("inf")))
        t_best = torch.minimum(t_best, tb)
    t = t_best + N(*t_best.shape) * 0.02
    keep = torch.isfinite(t_best) & (t > 0.5) & (t <= 120.0) & (R(*t.shape) >= 0.05)
    t, dx, dy, dz = t[keep], dx[keep], dy[keep], dz[keep]
    x, y, z = t * dx, t * dy, t * dz
    inten = (R(*t.shape) * 100).floor() / 100.0
    # reflection noise below the ground: steep downward angle, low intensity
    nref = int(5 + (R(1) * 16).item())
    rr = 4.0 + R(nref) * 6.0
    ra = R(nref) * 2 * math.pi
    rz = -2.6 - R(nref) * 1.4
    x = torch.cat([x, rr * torch.cos(ra)])
    y = torch.cat([y, rr * torch.sin(ra)])
    z = torch.cat([z, rz])
    inten = torch.cat([inten, R(nref) * 0.19])
    perm_tail = torch.randperm(x.shape[0], generator=gen, device=device)[:nref]
    # scatter the reflection points into the stream instead of leaving them at the end
    idx = torch.arange(x.shape[0], device=device)
    tail = idx[-nref:].clone()
    idx[-nref:] = idx[perm_tail]
    idx[perm_tail] = tail
    pts = torch.stack([x[idx], y[idx], z[idx], inten[idx]], dim=1).to(torch.float32)
    return pts.contiguous()


def make_batch(seed: int, first: int, count: int, sensor: str = "kitti64", device="cpu"):
    """Frames first..first+count-1 packed back to back: (points (total,4) float32, offsets int64 (count+1) on CPU)."""
    frames = [make_frame(seed, first + i, sensor, device) for i in range(count)]
    offs = [0]
    for fr in frames:
        offs.append(offs[-1] + fr.shape[0])
    pts = torch.cat(frames, dim=0) if frames else torch.zeros((0, 4), dtype=torch.float32, device=device)
    return pts, torch.tensor(offs, dtype=torch.int64)
