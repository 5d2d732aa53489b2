"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's SemanticKITTI window assembly, prediction
writer and checkpoint naming.  Nothing under temporal_latticenet_b200/ may import this.

Follows /root/reference/dataloader/kitti_dataloader.py:
  :100-114  window indices (frames_per_seq, cloud_scope, clamp at 0)
  :122,160-171  velo -> world -> first scan of the window -> rotation_matrix(-90, "x"), divide by w
  :129-132  .bin wire format ([N,4] float32)      :281-291  .label wire format (uint16 pairs -> learning_map LUT)
  :205-256  calib.txt / poses.txt parsing (pose = Tr^-1 . P . Tr)
DataTransformer.py:19-31 (rotation_matrix via scipy), :88-91 (float64 -> float32 for the non-training splits),
test_ln.py:219-231 (uint32 labels written as decimal text, one per line), train_ln.py:248-249 (checkpoint name).
Pinned against the reference's OWN loader run in the build container: tests/golden/make_kitti_golden.py ->
tests/golden/kitti_window.npz (tests/test_kitti_io_cpu.py).
"""
import os

import numpy as np
from numpy.linalg import inv


def parse_calibration(filename):
    calib = {}
    with open(filename) as f:
        for line in f:
            key, content = line.strip().split(":")
            values = [float(v) for v in content.strip().split()]
            pose = np.zeros((4, 4))
            pose[0, 0:4], pose[1, 0:4], pose[2, 0:4], pose[3, 3] = values[0:4], values[4:8], values[8:12], 1.0
            calib[key] = pose
    return calib


def parse_poses(filename, calibration):
    Tr = calibration["Tr"]
    Tr_inv = inv(Tr)
    poses = []
    with open(filename) as f:
        for line in f:
            values = [float(v) for v in line.strip().split()]
            pose = np.zeros((4, 4))
            pose[0, 0:4], pose[1, 0:4], pose[2, 0:4], pose[3, 3] = values[0:4], values[4:8], values[8:12], 1.0
            poses.append(np.matmul(Tr_inv, np.matmul(pose, Tr)))
    return poses


def rotation_matrix_x(angle_deg):
    from scipy.spatial.transform import Rotation as R
    T = np.identity(4)
    T[:3, :3] = R.from_euler("X", angle_deg, degrees=True).as_matrix()
    return T


def window_indices(index, frames, scope):
    return np.maximum((np.arange(frames) - (frames - 1)) * scope + index, 0)


def load_label(path, remap_lut):
    npz = np.fromfile(path, dtype=np.uint16)
    labels = (npz[0::2].reshape(len(npz) // 2, 1)).astype(np.int16)
    return np.squeeze(remap_lut[labels], axis=1)


def assemble_window(data_dir, seq, index, frames, scope, remap_lut=None, cap_distance=-1, min_distance=-1):
    """-> [(positions [N,3] float32, reflectance [N,1] float32, labels [N] or None)] for the window ending at scan `index`;
    cap_distance / min_distance: the training-time range crop of kitti_dataloader.py:142-154"""
    sdir = os.path.join(data_dir, "sequences", "%02d" % seq)
    poses = parse_poses(os.path.join(sdir, "poses.txt"), parse_calibration(os.path.join(sdir, "calib.txt")))
    idxs = window_indices(index, frames, scope)
    first = poses[idxs[0]]
    out = []
    for idx in idxs:
        raw = np.fromfile(os.path.join(sdir, "velodyne", "%06d.bin" % idx), dtype=np.float32).reshape(-1, 4).transpose()
        refl, xyz = raw[3, :], raw[0:3, :]
        lab = None
        if remap_lut is not None:
            lab = load_label(os.path.join(sdir, "labels", "%06d.label" % idx), remap_lut)
        if cap_distance >= 0:
            mask = np.linalg.norm(xyz, axis=0) < cap_distance
            xyz, refl = xyz[:, mask], refl[mask]
            lab = lab[mask] if lab is not None else None
        if min_distance >= 0:
            mask = np.linalg.norm(xyz, axis=0) > min_distance
            xyz, refl = xyz[:, mask], refl[mask]
            lab = lab[mask] if lab is not None else None
        hom = np.ones((4, xyz.shape[1]))
        hom[0:3, :] = xyz
        world = np.matmul(poses[idx], hom)
        world = np.matmul(np.linalg.inv(first), world)
        ros = np.matmul(rotation_matrix_x(-90), world)
        scan = (ros[0:3, :] / ros[3, :]).transpose()
        out.append((scan.astype(np.float32), np.expand_dims(refl, 1).astype(np.float32), lab))
    return out


def write_prediction(path, labels):
    """test_ln.py:219-231: the binary uint32 dump is immediately overwritten by decimal text, one label per line"""
    l_pred = np.asarray(labels).reshape(-1).astype(np.uint32)
    l_pred.tofile(path)
    with open(path, "w") as f:
        for i in range(l_pred.shape[0]):
            f.write(str(l_pred[i]) + "\n")


def checkpoint_name(date_time, moving, dataset_name, values_mode, sigma_0, rnn_modules, accumulate, frames, scope, epoch):
    return "{}_{}_{}_{}_sigma{}_type{}_frames{}_scope{}_epoch{}".format(
        date_time, "multi" if moving is True else "single", "Kitti" if dataset_name == "semantickitti" else "Paris",
        "Ref" if values_mode == "reflectance" else "xyz", str(sigma_0)[0:3], "-".join(rnn_modules) if not accumulate else "ACCUM",
        frames, scope, epoch)
