"""Golden fixtures = outputs of the reference's own serialized graphs (tools/make_golden.py, oracle/graph_exec.py).

CPU: the restated oracle must reproduce them (this is what pins the oracle).
GPU: the CUDA path must reproduce them through the C ABI."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import yamnet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))


def test_fixture_inventory():
    man = json.load(open(os.path.join(GOLD, "MANIFEST.json")))
    assert sorted(c["name"] for c in man["cases"]) == CASES and len(CASES) >= 6
    # the fixtures come from executing every node class of the reference graph
    ops = man["cases"][0]["graph_ops"]
    for op in ("RFFT", "ComplexAbs", "MatMul", "Log", "Conv2D", "DepthwiseConv2dNative", "FusedBatchNormV3", "Relu",
               "Mean", "GatherV2", "Pad", "Ceil"):
        assert ops.get(op, 0) > 0, op
    assert ops["Conv2D"] == 14 and ops["DepthwiseConv2dNative"] == 13 and ops["FusedBatchNormV3"] == 27


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_graph(case, yamnet_variables, mel, head):
    g = np.load(os.path.join(GOLD, case + ".npz"))
    hop = int(g["hop_frames"])
    act, emb = O.predict(g["samples"], yamnet_variables, mel, head[0], head[1], hop, return_embeddings=True)
    assert emb.shape == g["embeddings"].shape and act.shape == g["activations"].shape
    assert act.shape[0] == O.frame_counts(len(g["samples"]), hop)[2]          # frame indexing: exact
    e_err = np.abs(emb - g["embeddings"]).max() / max(np.abs(g["embeddings"]).max(), 1e-9)
    a_err = np.abs(act - g["activations"]).max()
    assert e_err <= 5e-6, e_err
    assert a_err <= 2e-5, a_err
    assert np.array_equal(np.round(act, 2) != np.round(g["activations"], 2),
                          np.zeros_like(act, dtype=bool)) or a_err <= 2e-5


@pytest.mark.skipif(not os.path.exists("/root/reference/embedders"), reason="reference checkout not present")
def test_fixtures_are_current(yamnet_variables):
    """Re-run the graph interpreter on one case: committed fixtures must match what the reference graph gives now."""
    from oracle import graph_exec as G
    g = np.load(os.path.join(GOLD, "whole_5s.npz"))
    emb, gf = G.run_yamnet_graph("/root/reference", g["samples"], yamnet_variables, "wholehop")
    act, _ = G.run_head_graph("/root/reference", emb)
    assert np.array_equal(emb, g["embeddings"]) and np.array_equal(act, g["activations"])
    # halfhop graph: same mel constant, hop constants 7680, different output key
    emb2, _ = G.run_yamnet_graph("/root/reference", g["samples"], yamnet_variables, "halfhop")
    assert emb2.shape[0] == O.frame_counts(len(g["samples"]), 48)[2]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("precision", ["fp16x3", "fp32"])
def test_cuda_reproduces_reference_graph(case, precision, engines):
    g = np.load(os.path.join(GOLD, case + ".npz"))
    hop = int(g["hop_frames"])
    e = engines(precision, early_patches=16, late_patches=48)
    act, emb = e.predict(g["samples"], hop, want_embeddings=True)
    assert act.shape == g["activations"].shape                               # frame indexing: exact
    a_err = float(np.abs(act - g["activations"]).max())
    e_err = float(np.abs(emb - g["embeddings"]).max() / np.abs(g["embeddings"]).max())
    assert a_err <= 1e-3, a_err                                               # north-star tolerance
    assert e_err <= 1e-4, e_err
    thr = float(np.median(g["activations"][:, 8]))
    near = np.abs(g["activations"][:, 8] - thr) <= 1e-3
    assert not ((act[:, 8] > thr) != (g["activations"][:, 8] > thr))[~near].any()
