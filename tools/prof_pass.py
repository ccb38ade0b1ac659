#!/usr/bin/env python
"""One device-resident pass over a 1-hour recording (for ncu): python tools/prof_pass.py [precision] [passes] [hop]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("BUZZ_B200_ALLOW_SYNTHETIC", "1")
import numpy as np
import torch
import __graft_entry__ as g
g.build()
from buzzdetect_b200 import capi
from oracle import yamnet_oracle as O

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hop = int(sys.argv[3]) if len(sys.argv) > 3 else 96
n = 3600 * 16000
x = torch.from_numpy(np.tile(O.synth_audio(60 * 16000, seed=1), 60)).cuda()
eng = capi.Engine(device=0, precision=prec, use_graph=False)
P = capi.frames_for(n, hop)[2]
act = torch.empty((P, eng.n_classes), dtype=torch.float32, device="cuda")
for _ in range(passes):
    eng.predict_device_ptr(x.data_ptr(), n, hop, act.data_ptr())
torch.cuda.synchronize()
prof = eng.profile_device_ptr(x.data_ptr(), n, hop)
print({k: round(v["ms"], 4) for k, v in prof.items() if k != "layers"})
print({k: round(v["pw_ms"], 4) for k, v in prof["layers"].items()})
eng.close()
