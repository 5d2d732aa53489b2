"""Device-side window assembly (csrc/ltn_io.cu through the C ABI) against the reference loader's own output on the tiny
fixture sequence and against a float64 numpy evaluation at full scan size.  Tolerance: the transform runs in float64 on
the device in the reference's order of operations and is rounded to float32 once; numpy's dgemm may associate the 4-term
dot products differently, so positions must be EQUAL for >= 99.9 % of the coordinates and within one float32 ulp for the
rest; the reflectance column is a copy (bit-exact)."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu
TINY = os.path.join(GOLDEN, "kitti_tiny")


def _ulp_check(got, want):
    assert got.shape == want.shape and got.dtype == np.float32
    same = got == want
    assert same.mean() >= 0.999, same.mean()
    ulp = np.spacing(np.abs(want).astype(np.float32))
    assert np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= ulp.astype(np.float64))


def test_window_from_wire_format_matches_reference_loader():
    from temporal_latticenet_b200.kitti_io import KittiSequence
    g = np.load(os.path.join(GOLDEN, "kitti_window.npz"))
    seq = KittiSequence(TINY, 8, frames=4, scope=3, device="cuda:0")
    frames = seq.window(7)
    assert len(frames) == 4
    for i, (pos, val) in enumerate(frames):
        _ulp_check(pos.cpu().numpy(), g["scan_%d" % i])
        assert np.array_equal(val.cpu().numpy(), g["feature_%d" % i])


def test_full_size_scan_transform_and_round_trip():
    from temporal_latticenet_b200 import kitti_io as P
    rng = np.random.default_rng(4)
    n = 125000
    raw = np.concatenate([rng.uniform(-80, 80, size=(n, 2)), rng.uniform(-3, 5, size=(n, 1)), rng.uniform(0, 1, size=(n, 1))], 1).astype(np.float32)
    yaw = 0.3
    pose = np.array([[np.cos(yaw), 0, np.sin(yaw), 12.5], [0, 1, 0, -0.4], [-np.sin(yaw), 0, np.cos(yaw), 230.0], [0, 0, 0, 1.0]])
    first = np.array([[np.cos(0.1), 0, np.sin(0.1), 10.0], [0, 1, 0, -0.3], [-np.sin(0.1), 0, np.cos(0.1), 221.0], [0, 0, 0, 1.0]])
    mats = np.stack([pose, np.linalg.inv(first), P.rotation_x(-90)])
    pos, val = P.assemble_scan(torch.from_numpy(raw).cuda(), mats)
    hom = np.ones((4, n)); hom[:3] = raw[:, :3].T.astype(np.float64)
    w = mats[2] @ (mats[1] @ (mats[0] @ hom))
    _ulp_check(pos.cpu().numpy(), (w[:3] / w[3]).T.astype(np.float32))
    assert np.array_equal(val.cpu().numpy()[:, 0], raw[:, 3])
    # round trip through the inverse chain: back to the sensor frame within float32 rounding of the intermediate
    back, _ = P.assemble_scan(torch.cat([pos, val], 1).contiguous(), np.stack([np.linalg.inv(mats[2]), first, np.linalg.inv(pose)]))
    assert np.abs(back.cpu().numpy() - raw[:, :3]).max() < 5e-5
    # empty scan
    p0, v0 = P.assemble_scan(torch.zeros(0, 4, device="cuda"), mats)
    assert p0.shape == (0, 3) and v0.shape == (0, 1)
