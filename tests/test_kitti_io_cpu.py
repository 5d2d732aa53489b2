"""SemanticKITTI wire formats / window assembly, CPU side: the numpy oracle against what the REFERENCE's own loader
returned on the tiny fixture sequence (tests/golden/make_kitti_golden.py), and the product's host-side functions
(parsing, indices, label remap, prediction files, checkpoint names) against the oracle."""
import os

import numpy as np
import yaml

from tests.helpers import GOLDEN, REPO

TINY = os.path.join(GOLDEN, "kitti_tiny")


def _golden():
    return np.load(os.path.join(GOLDEN, "kitti_window.npz"))


def test_oracle_window_matches_reference_loader_bit_exact():
    from oracle import kitti_oracle as K
    g = _golden()
    frames = K.assemble_window(TINY, 8, 7, frames=4, scope=3, remap_lut=g["remap_lut"])
    assert len(frames) == int(g["nr_frames"]) == 4
    for i, (pos, refl, lab) in enumerate(frames):
        assert np.array_equal(pos, g["scan_%d" % i])          # float64 pipeline rounded to float32: identical bits
        assert np.array_equal(refl, g["feature_%d" % i])
        assert np.array_equal(lab, g["label_%d" % i])
    assert [str(g["path_%d" % i]) for i in range(4)] == ["sequences/08/velodyne/%06d.bin" % k for k in (0, 1, 4, 7)]


def test_host_side_parsing_indices_labels_match_oracle():
    from oracle import kitti_oracle as K
    from temporal_latticenet_b200 import kitti_io as P
    sdir = os.path.join(TINY, "sequences", "08")
    calib_o, calib_p = K.parse_calibration(os.path.join(sdir, "calib.txt")), P.parse_calibration(os.path.join(sdir, "calib.txt"))
    assert sorted(calib_o) == sorted(calib_p) and all(np.array_equal(calib_o[k], calib_p[k]) for k in calib_o)
    po, pp = K.parse_poses(os.path.join(sdir, "poses.txt"), calib_o), P.parse_poses(os.path.join(sdir, "poses.txt"), calib_p)
    assert len(po) == len(pp) == 8 and all(np.array_equal(a, b) for a, b in zip(po, pp))
    for index, frames, scope in ((7, 4, 3), (0, 4, 3), (2, 4, 3), (100, 4, 3), (5, 1, 3), (9, 3, 1)):
        assert np.array_equal(K.window_indices(index, frames, scope), P.window_indices(index, frames, scope))
    assert np.allclose(K.rotation_matrix_x(-90), P.rotation_x(-90), rtol=0, atol=1e-15)
    with open(os.path.join(REPO, "configs", "semantic_kitti_label_maps.yaml")) as f:
        lm = yaml.safe_load(f)["learning_map"]
    lut = P.remap_lut(lm)
    g = _golden()
    assert np.array_equal(lut, g["remap_lut"])
    for i, k in enumerate((0, 1, 4, 7)):
        path = os.path.join(sdir, "labels", "%06d.label" % k)
        assert np.array_equal(P.load_labels(path, lut), g["label_%d" % i])
        assert np.array_equal(P.load_labels(path, lut), K.load_label(path, lut))


def test_prediction_files_and_checkpoint_names(tmp_path):
    from oracle import kitti_oracle as K
    from temporal_latticenet_b200 import kitti_io as P
    rng = np.random.default_rng(0)
    for labels in (rng.integers(0, 26, size=1000), np.zeros(0, dtype=np.int64), np.array([25])):
        a, b = os.path.join(str(tmp_path), "a.label"), os.path.join(str(tmp_path), "b.label")
        K.write_prediction(a, labels)
        P.write_prediction(b, labels)
        with open(a, "rb") as fa, open(b, "rb") as fb:
            assert fa.read() == fb.read()                        # byte-identical to what test_ln.py:219-231 leaves on disk
        assert np.array_equal(P.read_prediction(b), np.asarray(labels).astype(np.uint32))
    args = ("18102026_1200", True, "semantickitti", "reflectance", "0.6 3", ["gru", "gru", "aflow", "gru"], False, 4, 3, 7)
    assert P.checkpoint_name(*args) == K.checkpoint_name(*args) == "18102026_1200_multi_Kitti_Ref_sigma0.6_typegru-gru-aflow-gru_frames4_scope3_epoch7"
    args = ("01012027_0000", False, "parislille", "none", "0.9 3", ["none"] * 4, True, 1, 1, 0)
    assert P.checkpoint_name(*args) == K.checkpoint_name(*args) == "01012027_0000_single_Paris_xyz_sigma0.9_typeACCUM_frames1_scope1_epoch0"


def test_training_range_crop_matches_reference_loader():
    """kitti_dataloader.py:142-154 (cap_distance / min_distance, training split only): the oracle against the reference
    loader's train split run with every random augmentation off, and the product's keep-mask (float32 norm formed exactly
    like np.linalg.norm forms it) against the oracle's, bit for bit."""
    import torch
    from oracle import kitti_oracle as K
    from temporal_latticenet_b200 import kitti_io as P
    g = _golden()
    frames = K.assemble_window(TINY, 8, 7, frames=4, scope=3, remap_lut=g["remap_lut"], cap_distance=30, min_distance=8)
    assert [f[0].shape[0] for f in frames] == g["crop_len_seq"].tolist()
    for i, (pos, refl, lab) in enumerate(frames):
        assert np.array_equal(pos, g["crop_scan_%d" % i])
        assert np.array_equal(refl, g["crop_feature_%d" % i])
        assert np.array_equal(lab, g["crop_label_%d" % i])
    sdir = os.path.join(TINY, "sequences", "08")
    for k in (0, 1, 4, 7):
        raw = np.fromfile(os.path.join(sdir, "velodyne", "%06d.bin" % k), dtype=np.float32).reshape(-1, 4)
        length = np.linalg.norm(raw[:, :3].T, axis=0)
        for cap, mn in ((30, 8), (30, -1), (-1, 8), (12.5, 12.4)):
            want = np.ones(raw.shape[0], bool)
            if cap >= 0:
                want &= length < cap
            if mn >= 0:
                want &= length > mn
            got = P.range_mask(torch.from_numpy(raw), cap, mn)
            assert np.array_equal(got.numpy(), want)
    assert P.range_mask(torch.zeros(5, 4), -1, -1) is None
