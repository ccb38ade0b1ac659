#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the reference's serialized graphs (oracle/graph_exec.py).

Needs /root/reference, so it runs in the build container only; the fixtures travel to the GPU box.

    python tools/make_golden.py

Each fixture: seeded input audio (stored as its generator arguments + a float32 copy for short clips), the hop, the
embeddings the reference graph produces with the (synthetic, seeded) YAMNet weights, the activations its real head
graph produces, and a few intermediate tensors.  tests/test_golden.py checks the restated oracle against these on
CPU; tests/test_gpu_parity.py checks the CUDA path against them on the B200.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from buzzdetect_b200 import weights as W          # noqa: E402
from oracle import graph_exec as G                # noqa: E402
from oracle import yamnet_oracle as O             # noqa: E402

REF = os.environ.get("BUZZ_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, n_samples, seed, hop
    ("whole_5s", 16000 * 5 + 123, 11, "wholehop"),
    ("half_5s", 16000 * 5 + 123, 11, "halfhop"),
    ("short_0p3s", 4800, 12, "wholehop"),          # shorter than one patch: graph zero-pads to 15600
    ("exact_2patch", 15600 + 15360, 13, "wholehop"),
    ("ragged_12s", 16000 * 12 + 5, 2, "halfhop"),
    ("empty", 0, 0, "wholehop"),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    variables = W.synthetic_yamnet()
    manifest = {"weights": f"synthetic seed {W.SYNTH_SEED}", "cases": []}
    for name, n, seed, hop in CASES:
        x = O.synth_audio(max(n, 1), seed=seed)[:n]
        emb, g = G.run_yamnet_graph(REF, x, variables, hop)
        act, gh = G.run_head_graph(REF, emb)
        assert abs(g.last_bn_epsilon - 1e-4) < 1e-9
        np.savez_compressed(os.path.join(OUT, name + ".npz"), samples=x.astype(np.float32),
                            embeddings=emb.astype(np.float32), activations=act.astype(np.float32),
                            hop_frames=np.int32(96 if hop == "wholehop" else 48))
        manifest["cases"].append({"name": name, "n": n, "seed": seed, "hop": hop, "patches": int(emb.shape[0]),
                                  "graph_ops": dict(sorted(g.ops_run.items()))})
        print(name, emb.shape, act.shape, float(act.min()) if act.size else None, float(act.max()) if act.size else None)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
