"""Driver for the ncu launch list: cfg5 Yule-Walker (generic LWR) + transfer stage, 8 windows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hyperscanning_signal_analysis_b200 import _lib, synth
lib = _lib.load()
m, n, trials, p, nw, nf = 128, 512, 100, 15, 8, 512
ep = synth.cfg5_epochs(n_windows=nw)
x = torch.from_numpy(np.ascontiguousarray(ep.transpose(0, 3, 1, 2))).cuda()
offs = (torch.arange(nw * trials, dtype=torch.int64, device="cuda") * (m * n)).contiguous()
R = torch.empty((nw, p + 1, m, m), dtype=torch.float64, device="cuda")
A = torch.empty((nw, m, m, p), dtype=torch.float64, device="cuda")
V = torch.empty((nw, m, m), dtype=torch.float64, device="cuda")
ff = torch.empty((nw, m, m, nf), dtype=torch.float64, device="cuda")
st = torch.zeros(nw, dtype=torch.int32, device="cuda")
fr = torch.linspace(0, 128, nf + 1, dtype=torch.float64, device="cuda")[:nf].contiguous()
sp = torch.cuda.current_stream().cuda_stream
yws = torch.empty(lib.hs_yw_ws_bytes(nw, m, p), dtype=torch.uint8, device="cuda")
tws = torch.empty(lib.hs_transfer_ws_bytes(nw, m, p, nf), dtype=torch.uint8, device="cuda")
for _ in range(2):
    _lib.check(lib.hs_lagcov_f64(x.data_ptr(), offs.data_ptr(), n, nw, trials, m, n, p, R.data_ptr(), sp), "k3")
    _lib.check(lib.hs_yw_solve_f64(R.data_ptr(), nw, m, p, A.data_ptr(), V.data_ptr(), None, st.data_ptr(), yws.data_ptr(), sp), "k4")
    _lib.check(lib.hs_transfer_dtf_f64(A.data_ptr(), fr.data_ptr(), nf, 256.0, nw, m, p, None, None, None, ff.data_ptr(), st.data_ptr(), tws.data_ptr(), sp), "k5")
torch.cuda.synchronize()
print("ok", int(st.max()), float((ff.sum(dim=(2, 3)) - 1).abs().max()))
