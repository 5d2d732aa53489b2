"""SemanticKITTI wire formats and device-side window assembly.

Mirrors what the reference's loader and drivers do around the hot path, with the per-point arithmetic on the GPU:
  * dataloader/kitti_dataloader.py:100-114  window indices; :205-256 calib / poses; :129-132 .bin; :281-291 .label
  * :122,160-171  velo -> world -> first scan of the window -> -90 degrees about x  (csrc/ltn_io.cu, float64 on the device)
  * test_ln.py:219-231  prediction files (decimal text, one uint32 label per line)
  * train_ln.py:248-254 checkpoint names; test_ln.py:169-185 checkpoint loading protocol
At several hundred windows per second and GPU the numpy loader (4 matmuls over 120k points per scan on one core) is
the bottleneck; here a scan costs one pinned host read of the file bytes, one H2D copy and one kernel.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib


def parse_calibration(filename):
    """calib.txt -> {key: 4x4 float64} (kitti_dataloader.py:205-230)"""
    calib = {}
    with open(filename) as f:
        for line in f:
            if ":" not in line:
                continue
            key, content = line.strip().split(":")
            v = [float(x) for x in content.strip().split()]
            m = np.zeros((4, 4))
            m[0, :], m[1, :], m[2, :], m[3, 3] = v[0:4], v[4:8], v[8:12], 1.0
            calib[key] = m
    return calib


def parse_poses(filename, calibration):
    """poses.txt -> list of velo-to-world 4x4 float64, Tr^-1 . P . Tr (kitti_dataloader.py:232-256)"""
    Tr = calibration["Tr"]
    Tr_inv = np.linalg.inv(Tr)
    poses = []
    with open(filename) as f:
        for line in f:
            v = [float(x) for x in line.strip().split()]
            if len(v) < 12:
                continue
            m = np.zeros((4, 4))
            m[0, :], m[1, :], m[2, :], m[3, 3] = v[0:4], v[4:8], v[8:12], 1.0
            poses.append(np.matmul(Tr_inv, np.matmul(m, Tr)))
    return poses


def rotation_x(angle_deg):
    """DataTransformer.py:19-31 rotation_matrix(angle, "x")"""
    a = math.radians(angle_deg)
    c, s = math.cos(a), math.sin(a)
    T = np.identity(4)
    T[1, 1], T[1, 2], T[2, 1], T[2, 2] = c, -s, s, c
    return T


def window_indices(index, frames, scope):
    """scan numbers of the window ending at `index` (kitti_dataloader.py:100-114): clamped at the sequence start"""
    return np.maximum((np.arange(frames) - (frames - 1)) * scope + index, 0)


def remap_lut(learning_map):
    """learning_map {raw label: class} -> lookup table (kitti_dataloader.py:42-47, including its +100 slack)"""
    lut = np.zeros((max(learning_map.keys()) + 100), dtype=np.int32)
    lut[list(learning_map.keys())] = list(learning_map.values())
    return lut


def load_labels(path, lut):
    """.label: uint32 per point, lower 16 bits = label, upper 16 = instance id (kitti_dataloader.py:281-291)"""
    raw = np.fromfile(path, dtype=np.uint32)
    return lut[(raw & 0xFFFF).astype(np.int64)]


def write_prediction(path, labels):
    """what test_ln.py:219-231 leaves on disk: one decimal uint32 label per line (remap_semantic_labels.py reads it with
    np.fromfile(..., dtype=np.uint32, sep="\\n"))"""
    arr = torch.as_tensor(labels).reshape(-1).to("cpu").numpy().astype(np.uint32)
    with open(path, "w") as f:
        f.write("".join("%d\n" % int(x) for x in arr))


def read_prediction(path):
    return np.fromfile(path, dtype=np.uint32, sep="\n")


def checkpoint_name(date_time, include_moving_classes, dataset_name, values_mode, sigma_0, rnn_modules, accumulate_clouds,
                    frames_per_seq, cloud_scope, epoch):
    """train_ln.py:248-249"""
    return "{}_{}_{}_{}_sigma{}_type{}_frames{}_scope{}_epoch{}".format(
        date_time, "multi" if include_moving_classes is True else "single", "Kitti" if dataset_name == "semantickitti" else "Paris",
        "Ref" if values_mode == "reflectance" else "xyz", str(sigma_0)[0:3],
        "-".join(rnn_modules) if not accumulate_clouds else "ACCUM", frames_per_seq, cloud_scope, epoch)


class KittiSequence:
    """One sequence directory (sequences/XX with velodyne/, labels/, calib.txt, poses.txt) as a source of windows."""

    def __init__(self, data_dir, seq, frames=4, scope=3, learning_map=None, device=None):
        self.dir = os.path.join(data_dir, "sequences", "%02d" % int(seq))
        self.frames, self.scope = frames, scope
        self.poses = parse_poses(os.path.join(self.dir, "poses.txt"), parse_calibration(os.path.join(self.dir, "calib.txt")))
        self.lut = remap_lut(learning_map) if learning_map is not None else None
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._rx = rotation_x(-90.0)

    def scan_path(self, idx):
        return os.path.join(self.dir, "velodyne", "%06d.bin" % int(idx))

    def matrices(self, idx, first_idx):
        """the three float64 matrices of kitti_dataloader.py:160-167, in application order"""
        return np.ascontiguousarray(np.stack([self.poses[int(idx)], np.linalg.inv(self.poses[int(first_idx)]), self._rx]))

    def _staging(self, k, nr_floats):
        """reused pinned staging buffer of frame slot k (grown on demand): the .bin payload is read straight into it"""
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        buf = self._pinned.get(k)
        if buf is None or buf.numel() < nr_floats:
            buf = torch.empty(max(nr_floats, 4 * 140000), dtype=torch.float32).pin_memory()
            self._pinned[k] = buf
        return buf

    def read_raw(self, idx, slot=0):
        """the .bin payload as a pinned host tensor [N,4] float32: ONE read of the file into the slot's reused pinned
        buffer (no intermediate numpy array, no per-scan pinning, no per-point work on the host).  Valid until the slot is
        read into again."""
        path = self.scan_path(idx)
        nr_floats = os.path.getsize(path) // 4
        ev = getattr(self, "_copied", {}).get(slot)
        if ev is not None:
            ev.synchronize()        # the previous window's host-to-device copy out of this slot has completed
        buf = self._staging(slot, nr_floats)
        view = buf.numpy()[:nr_floats]
        with open(path, "rb") as f:
            got = f.readinto(memoryview(view).cast("B"))
        if got != 4 * nr_floats or nr_floats % 4:
            raise RuntimeError("%s is not a whole number of float32 x 4 points" % path)
        return buf[:nr_floats].view(-1, 4)

    def labels(self, idx):
        if self.lut is None:
            return None
        return torch.from_numpy(load_labels(os.path.join(self.dir, "labels", "%06d.label" % int(idx)), self.lut).astype(np.int64))

    def window(self, index, with_labels=False, cap_distance=-1.0, min_distance=-1.0, shuffle=False, generator=None):
        """-> [(positions [N,3], values [N,1])] on the device (+ labels of the last frame, on the HOST): the model's inputs for
        the window ending at scan `index`, every frame expressed in the first frame's coordinates (kitti_dataloader.py:119-171).
        Training-time options of the reference loader: cap_distance / min_distance >= 0 drop the points outside that range
        of the SENSOR (:142-154, before the pose transform, order preserved), shuffle permutes the points of every frame
        (:173-180; the same permutation for positions, values and labels).  `generator`, when given, must be a CUDA generator
        of this sequence's device (the permutation is drawn on the device).  Crop and shuffle act on the labels ON THE
        DEVICE too; they come back to the host once per window, so a frame costs one pinned read, one host-to-device copy
        and device work only."""
        _lib.require_cuda()
        if generator is not None and generator.device != self.device:
            raise RuntimeError("shuffle needs a generator on %s (torch.Generator(device=...)), got one on %s" % (self.device, generator.device))
        idxs = window_indices(index, self.frames, self.scope)
        out, last_labels = [], None
        for k, idx in enumerate(idxs):
            raw = self.read_raw(idx, slot=k).to(self.device, non_blocking=True)
            if not hasattr(self, "_copied"):
                self._copied = {}
            self._copied[k] = torch.cuda.Event()
            self._copied[k].record()
            lab = self.labels(idx) if (with_labels and k == len(idxs) - 1) else None
            if lab is not None:
                lab = lab.pin_memory().to(self.device, non_blocking=True)
            keep = range_mask(raw, cap_distance, min_distance)
            if keep is not None:
                raw = raw[keep]
                if lab is not None:
                    lab = lab[keep]
            if shuffle:
                perm = torch.randperm(raw.shape[0], device=raw.device, generator=generator)
                raw = raw[perm]
                if lab is not None:
                    lab = lab[perm]
            out.append(assemble_scan(raw, self.matrices(idx, idxs[0])))
            if lab is not None:
                last_labels = lab
        if with_labels:
            return out, (last_labels.cpu() if last_labels is not None else None)
        return out


def range_mask(raw, cap_distance=-1.0, min_distance=-1.0):
    """boolean keep-mask of kitti_dataloader.py:142-154 (None when both limits are off): distance of the raw sensor-frame
    point from the origin, in float32 exactly as np.linalg.norm forms it ((x*x + y*y) + z*z, each step rounded, then sqrt),
    strictly below cap_distance and strictly above min_distance.  Works on whatever device `raw` [N,4] lives on."""
    if cap_distance < 0 and min_distance < 0:
        return None
    x, y, z = raw[:, 0], raw[:, 1], raw[:, 2]
    length = torch.sqrt((x * x + y * y) + z * z)
    keep = torch.ones_like(length, dtype=torch.bool)
    if cap_distance >= 0:
        keep &= length < cap_distance
    if min_distance >= 0:
        keep &= length > min_distance
    return keep


def assemble_scan(raw_dev, mats):
    """raw_dev [N,4] float32 on the device, mats [k,4,4] float64 (host) applied in order -> (positions [N,3], values [N,1])"""
    if not raw_dev.is_cuda:
        raise RuntimeError("assemble_scan runs on the GPU; there is no CPU fallback")
    raw_dev = raw_dev.contiguous()
    n = raw_dev.shape[0]
    pos = torch.empty(n, 3, dtype=torch.float32, device=raw_dev.device)
    val = torch.empty(n, 1, dtype=torch.float32, device=raw_dev.device)
    m = np.ascontiguousarray(np.asarray(mats, dtype=np.float64).reshape(-1, 16))
    rc = _lib.load().ltn_assemble_scan(_lib.ptr(raw_dev), n, m.ctypes.data_as(ctypes.c_void_p), int(m.shape[0]), _lib.ptr(pos),
                                       _lib.ptr(val), _lib.stream())
    _lib.check(rc, "ltn_assemble_scan")
    return pos, val
