"""Small driver for ncu: one windowed MVAR+ffDTF batch (default 296 windows of 38 x 512, p=8, F=256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hyperscanning_signal_analysis_b200 import mtmvar, synth
n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = synth.dyad_eeg(seed=3, n_samples=256 * (n_win + 1), line_amp=0.0)
starts = np.arange(n_win) * 256
freqs = np.linspace(0, 128, 256, endpoint=False)
xd = torch.from_numpy(x).cuda()
for _ in range(reps):
    out = mtmvar.windowed_ffdtf(xd, starts, 512, freqs, 256.0, 8)
torch.cuda.synchronize()
print("ok", float(out[0].sum()))
