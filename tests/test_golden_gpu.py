"""GPU parity against the committed golden vectors (tests/golden/*.npz): outputs of the REFERENCE's
own seq_lattice/models.py::LNN_SEQ run on the CPU over the oracle (tests/golden/make_golden.py).

The CUDA path is driven the way the reference drives it (test_ln.py:149-185): a first window creates
the lazy parameters, the seeded state-dict is loaded BY NAME (so parameter names and shapes must
equal the reference's), reset_sequence(), fresh Lattice, the window is re-run.

Bar: vertex counts and keys bit-exact (numbering is deterministic, so no canonicalisation is even
needed; the sorted-key comparison is kept as the contract the north star states); features after
~40 stacked fp32 GEMM layers within FEATURE_TOL of the tensor's absolute maximum (fp32 accumulation
order differs between our kernels and the CPU BLAS that produced the goldens; the observed errors are
printed, the tolerance is ~4x the largest one observed on the B200).
"""
import json
import os

import hjson
import numpy as np
import pytest
import torch

from tests.helpers import CFG, GOLDEN, canonical_order, seeded_state

pytestmark = pytest.mark.gpu

FEATURE_TOL = 2e-4   # observed on the B200 (printed by the tests): 1.4e-5 ... 4.9e-5 over the six goldens; SURVEY 8(c) asks 1e-4-class for GEMM chains

GRAD_TOL = 1e-2       # observed: 3.1e-3

GOLDENS = ["gru_gru_aflow_gru", "lstm_cga_linear_maxpool", "maxpool_aflow_lstm_cga", "aflow_x4",
           "linear_none_none_gru", "single_frame"]


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        meta = json.load(f)
    return z, meta


def _cfg_for(meta, tmp_path):
    with open(CFG) as f:
        cfg = hjson.loads(f.read())
    cfg["model"]["rnn_modules"] = meta["rnn_modules"]
    cfg["model"]["sequence_learning"] = meta["sequence_learning"]
    cfg["loader_semantic_kitti"]["frames_per_seq"] = meta["frames"]
    path = os.path.join(str(tmp_path), meta["name"] + ".cfg")
    with open(path, "w") as f:
        f.write(hjson.dumps(cfg))
    return path


def _run_window(model, Lattice, cfg, frames, dev):
    lattice = Lattice.create(cfg, "lattice")
    outs = []
    for i, (p, v) in enumerate(frames):
        early = i != len(frames) - 1
        a, b, lattice = model(lattice, torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), early, False)
        outs.append((a, b, lattice.nr_lattice_vertices()))
    return outs, lattice


def _rel(a, b):
    return float(np.abs(a - b).max()) / (float(np.abs(b).max()) + 1e-12)


@pytest.mark.parametrize("name", GOLDENS)
def test_window_matches_reference_golden(name, tmp_path):
    from temporal_latticenet_b200.config import ConfigParser
    from temporal_latticenet_b200.lattice import Lattice, ModelParams
    from temporal_latticenet_b200.model import LatticeNetSeq

    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    z, meta = _load(name)
    cfg = _cfg_for(meta, tmp_path)
    frames = [(z["pos%d" % i], z["val%d" % i]) for i in range(meta["frames"])]
    with torch.no_grad():
        model = LatticeNetSeq(meta["nr_classes"], ModelParams.create(cfg), ConfigParser(cfg)).to(dev)
        model.train(False)
        _run_window(model, Lattice, cfg, frames, dev)  # creates the lazy parameters
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        ref_shapes = {k: tuple(v) for k, v in meta["shapes"].items()}
        assert shapes == ref_shapes, "state-dict names/shapes differ from the reference's"
        model.load_state_dict(seeded_state(shapes))
        model.reset_sequence()
        outs, lattice = _run_window(model, Lattice, cfg, frames, dev)

    for i in range(meta["frames"]):
        assert outs[i][2] == int(z["nv%d" % i]), "vertex count of frame %d" % i
    keys = lattice.hash_table.keys().cpu().numpy()
    gk = z["keys0"]
    assert np.array_equal(keys[canonical_order(keys)], gk[canonical_order(gk)])
    assert np.array_equal(keys, gk)
    for i in range(meta["frames"] - 1):
        if "out%d" % i in z.files:
            got = outs[i][0].cpu().numpy()
            assert got.shape == z["out%d" % i].shape
            e = _rel(got, z["out%d" % i])
            print("%s: late-fusion features of frame %d: err/absmax %.2e" % (name, i, e))
            assert e < FEATURE_TOL, "late-fusion features of frame %d: %g" % (i, e)
    logits = outs[-1][1].cpu().numpy()
    assert logits.shape == z["logits"].shape
    assert np.isfinite(logits).all() == np.isfinite(z["logits"]).all()
    err = _rel(logits, z["logits"])
    print("%s: logits err/absmax %.2e" % (name, err))
    assert err < FEATURE_TOL, "logits rel-to-absmax error %g" % err
    # class decisions: identical wherever the golden's top-2 margin exceeds the feature tolerance
    g = z["logits"]
    top2 = np.sort(g, 1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 4 * FEATURE_TOL * np.abs(g).max()
    assert np.array_equal(logits.argmax(1)[clear], g.argmax(1)[clear])


@pytest.mark.parametrize("path", ["unfused", "fused"])
def test_training_step_gradients_match_oracle_autograd(tmp_path, path, monkeypatch):
    """BPTT through a 2-frame window (train_ln.py:163-233): every parameter gradient of the CUDA path
    (conv transpose-as-gather, slice_classify / gather backward kernels, scatter_max, GroupNorm) equals
    torch autograd over the oracle's unfused CPU graph with the same seeded weights, and one AdamW
    step lowers the loss.

    Two forms of the training graph, two stated tolerances (relative to each gradient tensor's absolute maximum; the largest
    observed is printed):
    * "unfused" (LTN_TRAIN_UNFUSED=1): every layer as separate GroupNorm / im2row / fp32 GEMM kernels -- GRAD_TOL (observed 3e-3);
    * "fused" (default, funcs._FusedConv): every layer's forward AND its backward-data GEMM on the tensor cores (3-pass tf32
      split).  Each layer is within 2e-4 of the unfused composition (tests/test_conv_tc_gpu.py::test_fused_training_layer_...),
      i.e. within the convolution's stated bound 2e-5 * sum|a||w|, but a tensor-core accumulator truncates where an fp32 FMA
      chain rounds, and through ~80 layers of back-propagation (2 frames) those per-layer differences add up in the gradients
      whose terms cancel (GroupNorm biases, AFlow's two scalars).  Stated: every tensor within FUSED_GRAD_TOL of its abs-max
      AND the whole flattened gradient within FUSED_GRAD_L2 in relative L2 norm (what an optimizer step sees)."""
    FUSED_GRAD_TOL, FUSED_GRAD_L2 = 1e-1, 2e-2
    monkeypatch.setenv("LTN_TRAIN_UNFUSED", "1" if path == "unfused" else "0")
    from oracle import window_oracle as WO
    from temporal_latticenet_b200.config import ConfigParser
    from temporal_latticenet_b200.lattice import Lattice, ModelParams
    from temporal_latticenet_b200.lovasz import LovaszSoftmax
    from temporal_latticenet_b200.model import LatticeNetSeq

    dev = torch.device("cuda:0")
    z, meta = _load("linear_none_none_gru")
    meta = dict(meta, frames=2)
    cfg = _cfg_for(meta, tmp_path)
    frames = [(z["pos%d" % i], z["val%d" % i]) for i in range(2)]
    target = torch.from_numpy(np.random.default_rng(0).integers(0, 26, frames[-1][0].shape[0]))
    model = LatticeNetSeq(26, ModelParams.create(cfg), ConfigParser(cfg)).to(dev)
    model.train(True)
    lovasz, nll = LovaszSoftmax(ignore_index=0), torch.nn.NLLLoss(ignore_index=0)

    def window_loss():
        model.reset_sequence()
        lattice = Lattice.create(cfg, "lattice")
        for i, (p, v) in enumerate(frames):
            out, _, lattice = model(lattice, torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev), i != 1, True)
        t = target.to(dev)
        return 0.5 * lovasz(out, t) + 0.5 * nll(out, t)

    with torch.no_grad():
        window_loss()  # lazy parameters
    model.load_state_dict(seeded_state({k: tuple(v.shape) for k, v in model.state_dict().items()}))
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-3, amsgrad=True)
    loss0 = window_loss()
    opt.zero_grad()
    loss0.backward()

    # the same window under autograd on the oracle
    orc = WO.OracleWindowRunner(cfg, 26).materialise_parameters(frames)
    orc.model.train(True)
    orc.model.reset_sequence()
    ls = WO.Lattice.create(cfg, "lattice")
    for i, (p, v) in enumerate(frames):
        out, _, ls = orc.model(ls, torch.from_numpy(p), torch.from_numpy(v), i != 1)
    # the oracle side uses the oracle's OWN per-class restatement of the Lovasz-softmax loss (oracle/shims), not the product's
    from latticenet_py.lattice.lovasz_loss import LovaszSoftmax as OracleLovasz
    import latticenet_py
    assert os.path.join("oracle", "shims") in os.path.abspath(latticenet_py.__file__ or latticenet_py.__path__[0])
    oloss = 0.5 * OracleLovasz(ignore_index=0)(out, target) + 0.5 * nll(out, target)
    oloss.backward()
    assert abs(float(loss0) - float(oloss)) < 1e-4 * abs(float(oloss))
    ograds = {k: p.grad for k, p in orc.model.named_parameters()}
    bad, worst, num, den = [], 0.0, 0.0, 0.0
    tol = GRAD_TOL if path == "unfused" else FUSED_GRAD_TOL
    for k, p in model.named_parameters():
        if ".AFLOW.weight" in k:
            continue  # quirk Q4: created, never used
        og = ograds[k]
        assert p.grad is not None and og is not None, k
        scale = float(og.abs().max()) + 1e-12
        diff = p.grad.cpu() - og
        err = float(diff.abs().max()) / scale
        num += float((diff.double() ** 2).sum())
        den += float((og.double() ** 2).sum())
        worst = max(worst, err)
        if err > tol:
            bad.append((k, err))
    l2 = (num / den) ** 0.5
    print("%s: largest gradient error / absmax over all parameters: %.2e, relative L2 error of the whole gradient: %.2e" % (path, worst, l2))
    assert not bad, bad
    assert l2 < (GRAD_TOL if path == "unfused" else FUSED_GRAD_L2), l2
    opt.step()
    with torch.no_grad():
        loss1 = window_loss()
    assert float(loss1) < float(loss0)
