"""Time the loader's default FIR branch (dataloader.py:793-801: causal notch + 201-tap low-pass + 3049-tap high-pass, delay rolled out)
on 38 channels at 1024 Hz.  Usage: python tools/bench_fir_branch.py [seconds]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hyperscanning_signal_analysis_b200 import frontend, synth, dataloader
from scipy import signal
seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 600
fs = 1024.0
x = synth.dyad_eeg(seed=7, m=38, fs=fs, n_samples=int(seconds * fs), drift=True)
xd = torch.from_numpy(x).cuda()
b_n, a_n = signal.iirnotch(50.0, 30.0, fs)
bl = signal.firwin(201, 64.0, fs=fs)
bh = signal.firwin(3049, 1.0, fs=fs, pass_zero=False)
def run():
    return frontend.lfilter_fir_chain_dev(xd, (b_n, a_n), bl, bh, remove_dc=True)
run(); torch.cuda.synchronize()
ts = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); y = run(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
n = x.shape[1]
flops = 2.0 * 38 * n * (201 + 3049)
print(json.dumps({"seconds_of_signal": seconds, "ms": float(np.min(ts)), "fir_tflops": flops / (np.min(ts) * 1e-3) * 1e-12}))
