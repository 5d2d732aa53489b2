"""Seeded synthetic SemanticKITTI-shaped LiDAR windows.

There is no dataset in the container, so the bench / tests ray-cast a 64-beam spinning LiDAR over a
ground plane with random boxes and jittered "vegetation".  A window is `frames` scans of the same
world taken `scope` scans apart along the driving axis and expressed in the FIRST frame's sensor
coordinates, which is what dataloader/kitti_dataloader.py:122,159-167 does with the poses; the
y-up convention follows the -90 degree x-rotation at kitti_dataloader.py:166.

Host-side numpy only: this is input synthesis, not part of the timed hot path.
"""
import math

import numpy as np

NR_BEAMS = 64
NR_AZIMUTHS = 2000
SENSOR_HEIGHT = 1.73
MAX_RANGE = 80.0


class World:
    def __init__(self, seed=0, nr_boxes=200, extent=60.0, lane_half_width=15.0):
        rng = np.random.default_rng(seed)
        size = rng.uniform(1.5, 10.0, (nr_boxes, 2))
        height = rng.uniform(1.0, 8.0, (nr_boxes, 1))
        centre = rng.uniform(-extent, extent, (nr_boxes, 2))
        # keep a street-wide corridor around the ego free (pushes boxes outwards); with these
        # defaults a scan splats onto ~20k vertices at sigma 0.6, the size SURVEY.md 8d plans for
        near = np.abs(centre[:, 1]) < lane_half_width
        centre[near, 1] = np.sign(centre[near, 1] + 1e-9) * (lane_half_width + np.abs(centre[near, 1]))
        self.lo = np.concatenate([centre - size / 2, np.zeros((nr_boxes, 1))], 1)  # x, y, z(up)
        self.hi = np.concatenate([centre + size / 2, height], 1)
        self.veg = rng.random(nr_boxes) < 0.25
        self.seed = seed


def _ray_dirs():
    elev = np.deg2rad(np.linspace(-24.8, 2.0, NR_BEAMS))
    azim = np.linspace(-math.pi, math.pi, NR_AZIMUTHS, endpoint=False)
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    d = np.stack([ce * np.cos(azim)[None, :], ce * np.sin(azim)[None, :], np.broadcast_to(se, (NR_BEAMS, NR_AZIMUTHS))], -1)
    return d.reshape(-1, 3)


def scan(world, ego_x=0.0, seed=0, nr_points=None):
    """One scan taken at (ego_x, 0, SENSOR_HEIGHT), returned in the sensor frame of ego_x = 0
    rotated to the reference's y-up convention.  Returns positions [N,3] f32, reflectance [N,1] f32."""
    rng = np.random.default_rng(seed * 7919 + 13)
    d = _ray_dirs()
    o = np.array([ego_x, 0.0, SENSOR_HEIGHT])
    t_hit = np.full(d.shape[0], np.inf)
    is_veg = np.zeros(d.shape[0], bool)
    # ground plane z = 0
    down = d[:, 2] < -1e-6
    t_hit[down] = -o[2] / d[down, 2]
    # ray / axis-aligned-box slabs, vectorised over all boxes with torch CPU threads (fp32)
    import torch
    with torch.no_grad():
        dt = torch.from_numpy(d.astype(np.float32))
        dt = torch.where(dt.abs() < 1e-9, torch.full_like(dt, 1e-9), dt)
        inv = (1.0 / dt)[:, None, :]
        ot = torch.from_numpy(o.astype(np.float32))
        best = torch.from_numpy(t_hit.astype(np.float32))
        best_veg = torch.zeros(d.shape[0], dtype=torch.bool)
        veg_t = torch.from_numpy(world.veg)
        for b0 in range(0, world.lo.shape[0], 64):
            lo = torch.from_numpy(world.lo[b0:b0 + 64].astype(np.float32))[None]
            hi = torch.from_numpy(world.hi[b0:b0 + 64].astype(np.float32))[None]
            t1 = (lo - ot) * inv
            t2 = (hi - ot) * inv
            tn = torch.minimum(t1, t2).amax(-1)
            tf = torch.maximum(t1, t2).amin(-1)
            tn = torch.where((tf >= tn) & (tn > 0.5), tn, torch.full_like(tn, float("inf")))
            tb, j = tn.min(1)
            closer = tb < best
            best = torch.where(closer, tb, best)
            best_veg = torch.where(closer, veg_t[b0:b0 + 64][j], best_veg)
        t_hit = best.numpy().astype(np.float64)
        is_veg = best_veg.numpy()
    keep = np.isfinite(t_hit) & (t_hit < MAX_RANGE)
    t = t_hit[keep] + rng.normal(0.0, 0.02, keep.sum())
    p = o[None] + d[keep] * t[:, None]
    veg = is_veg[keep]
    p[veg] += rng.normal(0.0, 0.6, (int(veg.sum()), 3))
    p[:, 2] -= SENSOR_HEIGHT  # sensor frame of the first pose
    # z-up (x fwd, y left, z up) -> y-up: rotate -90 deg about x  =>  (x, z, -y)
    out = np.stack([p[:, 0], p[:, 2], -p[:, 1]], 1).astype(np.float32)
    refl = rng.random((out.shape[0], 1)).astype(np.float32)
    if nr_points is not None and out.shape[0] > nr_points:
        sel = np.sort(rng.choice(out.shape[0], nr_points, replace=False))
        out, refl = out[sel], refl[sel]
    return out, refl


def window(seed=0, frames=4, scope=3, nr_points=None, speed_m_per_scan=0.9):
    """A `frames`-scan window: list of (positions, reflectance), first frame's coordinates."""
    w = World(seed)
    return [scan(w, ego_x=f * scope * speed_m_per_scan, seed=seed * 131 + f, nr_points=nr_points)
            for f in range(frames)]


def labels(nr_points, nr_classes=26, seed=0):
    """Categorical labels (loss/IoU harness only; the hot path never reads labels)."""
    rng = np.random.default_rng(seed + 5)
    return rng.integers(0, nr_classes, nr_points, dtype=np.int64)
