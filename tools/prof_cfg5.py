"""Driver for ncu: the cfg5 lag-covariance contraction (128 ch x 512 samples x 100 epochs, p = 15), 2 windows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hyperscanning_signal_analysis_b200 import _lib, synth
lib = _lib.load()
m, n, trials, p, nw = 128, 512, 100, 15, 2
ep = synth.cfg5_epochs(n_windows=nw)
x = torch.from_numpy(np.ascontiguousarray(ep.transpose(0, 3, 1, 2))).cuda()
offs = (torch.arange(nw * trials, dtype=torch.int64, device="cuda") * (m * n)).contiguous()
R = torch.empty((nw, p + 1, m, m), dtype=torch.float64, device="cuda")
for _ in range(2):
    _lib.check(lib.hs_lagcov_f64(x.data_ptr(), offs.data_ptr(), n, nw, trials, m, n, p, R.data_ptr(), torch.cuda.current_stream().cuda_stream), "k3")
torch.cuda.synchronize()
print("ok", float(R[0, 0].trace()))
