"""Generates the SemanticKITTI fixtures by running the REFERENCE's own loader (dataloader/kitti_dataloader.py, unmodified,
imported from /root/reference) on a tiny synthetic sequence.  Run in the build container only:
    python tests/golden/make_kitti_golden.py
Writes tests/golden/kitti_tiny/ (sequence 08: 8 scans x ~300 points, labels, calib.txt, poses.txt -- synthetic, seeded) and
tests/golden/kitti_window.npz (what SemanticKittiDataset("valid").__getitem__(7) returned: 4 frames, scope 3 -> scans 0,1,4,7).
`easypbr` (the viewer package the loader star-imports) is absent; an empty stub module stands in for it.
"""
import os
import sys
import types

import hjson
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REFERENCE = "/root/reference"
TINY = os.path.join(HERE, "kitti_tiny")


def make_dataset():
    rng = np.random.default_rng(8)
    sdir = os.path.join(TINY, "sequences", "08")
    os.makedirs(os.path.join(sdir, "velodyne"), exist_ok=True)
    os.makedirs(os.path.join(sdir, "labels"), exist_ok=True)
    Tr = np.array([[4.276802385584e-04, -9.999672484946e-01, -8.084491683471e-03, -1.198459927713e-02],
                   [-7.210626507497e-03, 8.081198471645e-03, -9.999413164504e-01, -5.403984729748e-02],
                   [9.999738645903e-01, 4.859485810390e-04, -7.206933692422e-03, -2.921968648686e-01]])
    with open(os.path.join(sdir, "calib.txt"), "w") as f:
        for k in ("P0", "P1", "P2", "P3"):
            f.write("%s: %s\n" % (k, " ".join("%.12e" % x for x in rng.normal(size=12))))
        f.write("Tr: %s\n" % " ".join("%.12e" % x for x in Tr.reshape(-1)))
    raw_labels = np.array([0, 1, 10, 11, 13, 15, 30, 40, 44, 48, 50, 51, 52, 60, 70, 71, 72, 80, 81, 99, 252, 253, 254, 255, 256, 257, 258, 259])
    with open(os.path.join(sdir, "poses.txt"), "w") as f:
        for i in range(8):
            yaw = 0.02 * i
            c, s = np.cos(yaw), np.sin(yaw)
            P = np.array([[c, 0, s, 0.05 * i], [0, 1, 0, -0.01 * i], [-s, 0, c, 0.9 * i]])   # camera frame: z forward
            f.write(" ".join("%.12e" % x for x in P.reshape(-1)) + "\n")
            n = 280 + 7 * i
            pts = np.concatenate([rng.uniform(-40, 40, size=(n, 2)), rng.uniform(-2, 3, size=(n, 1)), rng.uniform(0, 1, size=(n, 1))], 1)
            pts.astype(np.float32).tofile(os.path.join(sdir, "velodyne", "%06d.bin" % i))
            lab = rng.choice(raw_labels, size=n).astype(np.uint32) | (rng.integers(0, 500, size=n).astype(np.uint32) << 16)
            lab.tofile(os.path.join(sdir, "labels", "%06d.label" % i))


def main():
    make_dataset()
    sys.modules["easypbr"] = types.ModuleType("easypbr")          # viewer package, star-imported by the loader, unused here
    sys.path.insert(0, REFERENCE)
    sys.path.insert(0, os.path.join(REFERENCE, "dataloader"))
    from cfgParser import cfgParser                                # reference file
    from kitti_dataloader import SemanticKittiDataset              # reference file
    with open(os.path.join(REFERENCE, "seq_config", "lnn_eval_semantic_kitti.cfg")) as f:
        cfg = hjson.loads(f.read())
    lk = cfg["loader_semantic_kitti"]
    lk["dataset_path"] = TINY
    lk["yaml_config"] = os.path.join(REFERENCE, "seq_config", "semantic-kitti.yaml")
    lk["yaml_config_all"] = os.path.join(REFERENCE, "seq_config", "semantic-kitti-all.yaml")
    tmp_cfg = os.path.join("/tmp", "kitti_tiny.cfg")
    with open(tmp_cfg, "w") as f:
        f.write(hjson.dumps(cfg))
    ds = SemanticKittiDataset("valid", cfgParser(tmp_cfg), sequence_learning=True)
    scan_seq, feature_seq, label_seq, path_seq, len_seq = ds[7]
    out = {"nr_frames": np.int64(len(scan_seq)), "len_seq": np.asarray(len_seq), "remap_lut": ds.remap_lut}
    for i, (s, f_, l) in enumerate(zip(scan_seq, feature_seq, label_seq)):
        out["scan_%d" % i] = s.numpy()
        out["feature_%d" % i] = f_.numpy()
        out["label_%d" % i] = l.numpy()
        out["path_%d" % i] = np.asarray(os.path.relpath(path_seq[i], TINY))
    np.savez_compressed(os.path.join(HERE, "kitti_window.npz"), **out)
    # training split of the same loader: range crop on (cap 30 m, min 8 m), point shuffle and every augmentation off, so the
    # output is deterministic.  The train split starts with sequence 00: a copy of the tiny sequence under /tmp.
    import shutil
    tmp_root = "/tmp/kitti_tiny_train"
    shutil.rmtree(tmp_root, ignore_errors=True)
    shutil.copytree(os.path.join(TINY, "sequences", "08"), os.path.join(tmp_root, "sequences", "00"))
    lk["dataset_path"] = tmp_root
    lk["cap_distance"], lk["min_distance"], lk["shuffle_points"] = 30, 8, False
    tr = lk["transformer"]
    for k in ("random_translation_xyz_magnitude", "random_translation_xz_magnitude", "rotation_y_max_angle", "random_stretch_xyz_magnitude",
              "random_subsample_percentage", "chance_of_xyz_noise"):
        tr[k] = 0.0
    for k in ("random_mirror_x", "random_mirror_z", "random_rotation_90_degrees_y"):
        tr[k] = False
    with open(tmp_cfg, "w") as f:
        f.write(hjson.dumps(cfg))
    ds_train = SemanticKittiDataset("train", cfgParser(tmp_cfg), sequence_learning=True)
    scan_seq, feature_seq, label_seq, path_seq, len_seq = ds_train[7]
    for i, (s_, f_, l) in enumerate(zip(scan_seq, feature_seq, label_seq)):
        out["crop_scan_%d" % i] = s_.numpy()
        out["crop_feature_%d" % i] = f_.numpy()
        out["crop_label_%d" % i] = l.numpy()
    out["crop_len_seq"] = np.asarray(len_seq)
    np.savez_compressed(os.path.join(HERE, "kitti_window.npz"), **out)
    print("train split, cap 30 / min 8:", len_seq)
    # the dataset's label maps (facts of SemanticKITTI, read from the reference's yaml) for the product's configs/
    import yaml
    with open(lk["yaml_config_all"]) as f:
        data = yaml.safe_load(f)
    with open(os.path.join(REPO, "configs", "semantic_kitti_label_maps.yaml"), "w") as f:
        f.write("# SemanticKITTI label maps with the moving classes (26 training classes), extracted by tests/golden/make_kitti_golden.py\n")
        yaml.safe_dump({"learning_map": data["learning_map"], "learning_map_inv": data["learning_map_inv"]}, f, sort_keys=True)
    print("valid split:", int(out["nr_frames"]), "frames,", out["len_seq"].tolist(), "points")


if __name__ == "__main__":
    main()
