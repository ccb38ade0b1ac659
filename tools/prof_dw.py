"""Per-stage and per-depthwise-layer device times of one 1-hour pass (CUDA events around every launch): python tools/prof_dw.py"""
import os,sys
sys.path.insert(0,".")
os.environ["BUZZ_B200_ALLOW_SYNTHETIC"]="1"
import numpy as np, torch
from buzzdetect_b200 import capi
from oracle import yamnet_oracle as O
x = torch.from_numpy(np.tile(O.synth_audio(60 * 16000, seed=1), 60)).cuda()
eng = capi.Engine(device=0, precision="fp16x3", use_graph=False)
P = capi.frames_for(3600*16000, 96)[2]
act = torch.empty((P, eng.n_classes), dtype=torch.float32, device="cuda")
for _ in range(2): eng.predict_device_ptr(x.data_ptr(), 3600*16000, 96, act.data_ptr())
prof = eng.profile_device_ptr(x.data_ptr(), 3600*16000, 96)
print({k: round(v["ms"], 4) for k, v in prof.items() if k != "layers"})
print({k: round(v["dw_ms"],4) for k,v in prof["layers"].items() if v["dw_ms"]>0})
