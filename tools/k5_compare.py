"""Kernel-only time of the optimistic A(f)^-1 pass for both kernels (599 cfg2 windows), CUDA events inside the library."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hyperscanning_signal_analysis_b200 import _lib, synth
lib = _lib.load()
M, P, F, NW = 38, 8, 256, 599
x = synth.dyad_eeg(seed=1, n_samples=512 + 256 * (NW - 1), line_amp=0.0)
xd = torch.from_numpy(x).cuda()
starts = (torch.arange(NW, dtype=torch.int64, device="cuda") * 256)
T = x.shape[1]
R = torch.empty((NW, P + 1, M, M), dtype=torch.float64, device="cuda")
A = torch.empty((NW, M, M, P), dtype=torch.float64, device="cuda")
V = torch.empty((NW, M, M), dtype=torch.float64, device="cuda")
st = torch.zeros(NW, dtype=torch.int32, device="cuda")
out = torch.empty((NW, M, M, F), dtype=torch.float64, device="cuda")
fr = torch.linspace(0, 128, F + 1, dtype=torch.float64, device="cuda")[:F].contiguous()
sp = torch.cuda.current_stream().cuda_stream
lib.hs_lagcov_f64(xd.data_ptr(), starts.data_ptr(), T, NW, 1, M, 512, P, R.data_ptr(), sp)
yws = torch.empty(lib.hs_yw_ws_bytes(NW, M, P), dtype=torch.uint8, device="cuda")
lib.hs_yw_solve_f64(R.data_ptr(), NW, M, P, A.data_ptr(), V.data_ptr(), None, st.data_ptr(), yws.data_ptr(), sp)
tws = torch.empty(lib.hs_transfer_ws_bytes(NW, M, P, F), dtype=torch.uint8, device="cuda")
res = {}
ref = None
for which in (1, 4, 1, 4):
    _lib.check(lib.hs_transfer_set_kernel(which), "set")
    lib.hs_timing_enable(1)
    ts = []
    for i in range(8):
        _lib.check(lib.hs_transfer_dtf_f64(A.data_ptr(), fr.data_ptr(), F, 256.0, NW, M, P, None, None, None, out.data_ptr(), st.data_ptr(), tws.data_ptr(), sp), "k5")
        ms = C.c_double(0)
        lib.hs_timing_last_k5_ms(C.byref(ms))
        if i >= 3:
            ts.append(ms.value)
    lib.hs_timing_enable(0)
    off = lib.hs_transfer_ws_flag_offset(NW, M, P, F)
    flagged = int(tws[off:off + 4].view(torch.int32).item())
    res.setdefault(which, []).append({"ms": float(np.mean(ts)), "flagged": flagged, "rowsum_err": float((out.sum(dim=(2, 3)) - 1).abs().max())})
    if ref is None:
        ref = out.clone()
    else:
        res[which][-1]["max_rel_diff_vs_first"] = float(((out - ref).abs().max() / ref.abs().max()).item())
lib.hs_transfer_set_kernel(0)
print(json.dumps(res))
