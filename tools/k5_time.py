"""Time the K5 stage (hs_transfer_dtf_f64, dtf only) on the bench workload and report how many matrices
the optimistic pass flagged for the pivoted redo.  Usage: python tools/k5_time.py [n_win] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hyperscanning_signal_analysis_b200 import _lib
if os.environ.get('HS_LIB'): _lib.LIB_PATH = os.path.abspath(os.environ['HS_LIB'])      # a library variant built next to the product one
n_win_req = int(sys.argv[1]) if len(sys.argv) > 1 else 599
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
M, WIN, P, F, FS = bench.M, bench.WIN, bench.P, bench.F, bench.FS
y, starts, freqs = bench.make_task(20260101)
starts = starts[:n_win_req]
n_win = len(starts)
lib = _lib.load()
x_d = torch.from_numpy(y).cuda(); st_d = torch.from_numpy(starts).cuda(); fr_d = torch.from_numpy(freqs).cuda()
T = y.shape[1]
R = torch.empty((n_win, P + 1, M, M), dtype=torch.float64, device="cuda")
A = torch.empty((n_win, M, M, P), dtype=torch.float64, device="cuda")
V = torch.empty((n_win, M, M), dtype=torch.float64, device="cuda")
status = torch.zeros(n_win, dtype=torch.int32, device="cuda")
out = torch.empty((n_win, M, M, F), dtype=torch.float64, device="cuda")
yw_ws = torch.empty(lib.hs_yw_ws_bytes(n_win, M, P), dtype=torch.uint8, device="cuda")
tr_bytes = lib.hs_transfer_ws_bytes(n_win, M, P, F)
tr_ws = torch.zeros(tr_bytes, dtype=torch.uint8, device="cuda")
sp = torch.cuda.current_stream().cuda_stream
_lib.check(lib.hs_lagcov_f64(x_d.data_ptr(), st_d.data_ptr(), T, n_win, 1, M, WIN, P, R.data_ptr(), sp), "lagcov")
_lib.check(lib.hs_yw_solve_f64(R.data_ptr(), n_win, M, P, A.data_ptr(), V.data_ptr(), None, status.data_ptr(), yw_ws.data_ptr(), sp), "yw")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
import ctypes as C
lib.hs_timing_enable(1)
ks = []
ts = []
for i in range(reps + 2):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(lib.hs_transfer_dtf_f64(A.data_ptr(), fr_d.data_ptr(), F, FS, n_win, M, P, None, None, out.data_ptr(), None,
                                       status.data_ptr(), tr_ws.data_ptr(), sp), "transfer")
    b.record(); torch.cuda.synchronize()
    if i >= 2:
        ts.append(a.elapsed_time(b))
        kms = C.c_double(0.0)
        if lib.hs_timing_last_k5_ms(C.byref(kms)) == 0: ks.append(kms.value)
off = lib.hs_transfer_ws_flag_offset(n_win, M, P, F)
bad = int(tr_ws[off:off + 4].view(torch.int32).item())
fl = bench.flops_per_window()["transfer"] * n_win
print({"k5_kernel_ms": float(np.min(ks)) if ks else None, "n_win": n_win, "ms": float(np.mean(ts)), "min_ms": float(np.min(ts)), "tflops": fl / (np.min(ts) * 1e-3) * 1e-12,
       "flagged": bad, "status_max": int(status.max())})
