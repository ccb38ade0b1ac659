#!/usr/bin/env python
"""Extract the DATA the hot path needs from the reference tree into buzzdetect_b200/assets/.

Run in the build container (needs /root/reference; the GPU box has no copy of it):

    python tools/extract_assets.py [--reference /root/reference]

Nothing here is source code of the reference: the outputs are model constants and weights the
reference ships next to its plugins:

  mel_257x64.f32            Const_1 of embedders/yamnet_k2/models/yamnet_wholehop/saved_model.pb
                            (bit-identical in yamnet_halfhop)
  mel_yamnet_257x64.f32     Const_1 of embedders/yamnet/saved_model.pb (differs by <= 6.2e-6 in 20 entries)
  frontend_consts.json      Const / Const_3 / Const_5 ... (log offset, patch hop samples, min samples)
  head_kernel_1024x13.f32   models/model_general_v3/variables, layer_with_weights-0/kernel (CRC checked)
  head_bias_13.f32          ... layer_with_weights-0/bias (CRC checked)
  yamnet_tensor_table.json  name/shape/offset/size/masked-crc32c of the 108 YAMNet tensors from
                            embedders/yamnet/variables/variables.index (the .data blob itself is NOT
                            shipped in /root/reference -- see .MISSING_LARGE_BLOBS)
  config_model.json         models/model_general_v3/config_model.json (class list)
  metrics.csv               models/model_general_v3/tests/metrics.csv (threshold <-> precision table)
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tfbundle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "buzzdetect_b200", "assets"))
    args = ap.parse_args()
    R = args.reference
    out = os.path.abspath(args.out)
    os.makedirs(out, exist_ok=True)

    # ---- frontend constants from the three YAMNet graphs
    graphs = {
        "yamnet_k2/wholehop": "embedders/yamnet_k2/models/yamnet_wholehop/saved_model.pb",
        "yamnet_k2/halfhop": "embedders/yamnet_k2/models/yamnet_halfhop/saved_model.pb",
        "yamnet": "embedders/yamnet/saved_model.pb",
    }
    mel_ref = None
    consts_out = {}
    for key, rel in graphs.items():
        mg = tfbundle.read_meta_graphs(os.path.join(R, rel))[0]
        c = tfbundle.const_nodes(mg)
        mel = np.ascontiguousarray(c["Const_1"], dtype="<f4")
        assert mel.shape == (257, 64)
        if mel_ref is None:
            mel_ref = mel
        if key.startswith("yamnet_k2"):
            assert mel.tobytes() == mel_ref.tobytes(), f"mel constant differs in {key}"
        else:
            # embedders/yamnet/saved_model.pb was exported on another machine: 20 of its 461 non-zero
            # entries differ from the yamnet_k2 constant by <= 6.2e-6 (float32 linear_to_mel_weight_matrix
            # is recomputed by TF wherever the Keras-3 model is built).  Ship it separately.
            mel.tofile(os.path.join(out, "mel_yamnet_257x64.f32"))
            consts_out["mel_yamnet_sha256"] = hashlib.sha256(mel.tobytes()).hexdigest()
            consts_out["mel_yamnet_max_abs_diff_vs_k2"] = float(np.abs(mel - mel_ref).max())
        sig = mg.signature_def["serving_default"]
        consts_out[key] = {
            "tf_version": mg.meta_info_def.tensorflow_version,
            "log_offset": float(c["Const"]),
            "patch_hop_samples": int(c["Const_3"]),
            "min_samples": int(c["Const_5"]),
            "input_key": list(sig.inputs.keys())[0],
            "output_key": list(sig.outputs.keys())[0],
        }
    mel_ref.tofile(os.path.join(out, "mel_257x64.f32"))
    consts_out["mel_sha256"] = hashlib.sha256(mel_ref.tobytes()).hexdigest()
    consts_out["mel_nonzeros"] = int((mel_ref != 0).sum())
    with open(os.path.join(out, "frontend_consts.json"), "w") as f:
        json.dump(consts_out, f, indent=1, sort_keys=True)

    # ---- head weights (present in the reference), CRC-verified
    hdir = os.path.join(R, "models/model_general_v3")
    hidx = {e.name: e for e in tfbundle.read_bundle_index(os.path.join(hdir, "variables/variables.index"))}
    hdata = open(os.path.join(hdir, "variables/variables.data-00000-of-00001"), "rb").read()
    for short, name in (("head_kernel_1024x13", "layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE"),
                        ("head_bias_13", "layer_with_weights-0/bias/.ATTRIBUTES/VARIABLE_VALUE")):
        e = hidx[name]
        blob = hdata[e.offset:e.offset + e.size]
        assert tfbundle.masked_crc32c(blob) == e.crc32c, f"CRC mismatch for {name}"
        open(os.path.join(out, short + ".f32"), "wb").write(blob)
    shutil.copyfile(os.path.join(hdir, "config_model.json"), os.path.join(out, "config_model.json"))
    shutil.copyfile(os.path.join(hdir, "tests/metrics.csv"), os.path.join(out, "metrics.csv"))

    # ---- YAMNet tensor table (blob absent from the reference checkout)
    idx_paths = [
        "embedders/yamnet/variables/variables.index",
        "embedders/yamnet_k2/models/yamnet_wholehop/variables/variables.index",
        "embedders/yamnet_k2/models/yamnet_halfhop/variables/variables.index",
    ]
    blobs = [open(os.path.join(R, p), "rb").read() for p in idx_paths]
    assert blobs[0] == blobs[1] == blobs[2], "the three YAMNet variables.index files differ"
    ents = tfbundle.read_bundle_index(os.path.join(R, idx_paths[0]))
    table = [
        {"name": e.name.replace("/.ATTRIBUTES/VARIABLE_VALUE", ""), "shape": list(e.shape),
         "offset": e.offset, "size": e.size, "crc32c": e.crc32c}
        for e in ents if e.dtype == 1
    ]
    table.sort(key=lambda d: d["offset"])
    total = sum(d["size"] for d in table)
    with open(os.path.join(out, "yamnet_tensor_table.json"), "w") as f:
        json.dump({"data_file": "variables.data-00000-of-00001", "float_bytes": total,
                   "file_bytes": max(e.offset + e.size for e in ents), "tensors": table}, f, indent=1)
    print(f"wrote assets to {out}: mel, head, {len(table)} yamnet tensors ({total} B)")


if __name__ == "__main__":
    main()
