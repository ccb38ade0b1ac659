#!/usr/bin/env python
"""Build assets/synth_calibration.json: per-stage (input mean, centred pre-BN variance) so that the
SYNTHETIC YAMNet weights (buzzdetect_b200/weights.py:synthetic_yamnet) keep O(1) activations.

Build-time tool (uses the oracle's frontend); the product only reads the resulting 27x2 table."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from buzzdetect_b200 import weights as W          # noqa: E402
from oracle import yamnet_oracle as O             # noqa: E402


def main():
    x = O.synth_audio(16000 * 20, seed=123)
    mel = W.load_mel()
    lm = O.log_mel(O.pad_waveform(x, 96), mel)
    t = torch.from_numpy(O.patches_from_logmel(lm, 96)).double().unsqueeze(1)     # [P,1,96,64]
    calib = []
    names = W.layer_tensor_names()
    stage = 0
    for (kind, s, cin, cout, H, Wd), nm in zip(W.LAYERS, names):
        subs = ["conv"] if kind == "conv" else ["dw", "pw"]
        for sub in subs:
            mu_in = float(t.mean())
            trial = W.synthetic_yamnet(calib=calib + [(mu_in, 1.0)], upto=stage + 1)
            if sub == "conv":
                w = torch.from_numpy(trial[nm["w"]]).double().permute(3, 2, 0, 1)
                ph, pw = O._same_pad(t.shape[2], s), O._same_pad(t.shape[3], s)
                y = F.conv2d(F.pad(t, (pw[0], pw[1], ph[0], ph[1])), w, stride=s)
                bnp = nm["bn"]
            elif sub == "dw":
                w = torch.from_numpy(trial[nm["dw"]]).double().permute(2, 3, 0, 1)
                ph, pw = O._same_pad(t.shape[2], s), O._same_pad(t.shape[3], s)
                y = F.conv2d(F.pad(t, (pw[0], pw[1], ph[0], ph[1])), w, stride=s, groups=t.shape[1])
                bnp = nm["dw_bn"]
            else:
                w = torch.from_numpy(trial[nm["w"]]).double().permute(3, 2, 0, 1)
                y = F.conv2d(t, w)
                bnp = nm["bn"]
            mean = torch.from_numpy(trial[bnp + "/moving_mean"]).double().view(1, -1, 1, 1)
            v = float(((y - mean) ** 2).mean())
            calib.append((mu_in, v))
            final = W.synthetic_yamnet(calib=calib, upto=stage + 1)
            var = torch.from_numpy(final[bnp + "/moving_variance"]).double().view(1, -1, 1, 1)
            beta = torch.from_numpy(final[bnp + "/beta"]).double().view(1, -1, 1, 1)
            t = torch.relu((y - mean) * torch.rsqrt(var + W.BN_EPS) + beta)
            print(f"stage {stage:2d} {nm.get('w')} {sub}: mu_in={mu_in:.4f} var={v:.4f} out mean={float(t.mean()):.3f} "
                  f"max={float(t.max()):.2f} dead={float((t == 0).double().mean()):.2f}")
            stage += 1
    with open(os.path.join(W.ASSETS, "synth_calibration.json"), "w") as f:
        json.dump({"seed": W.SYNTH_SEED, "audio": "oracle.synth_audio(320000, seed=123)", "stages": calib}, f, indent=1)


if __name__ == "__main__":
    main()
