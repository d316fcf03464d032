"""Accuracy of the Yule-Walker stage against the reference goldens: norm-wise max|a-b|/max|b| of A, V and of the ffDTF / dDTF
computed from them.  Usage: python tools/k4_accuracy.py [path/to/libhs_b200.so]"""
import os, sys, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hyperscanning_signal_analysis_b200 import _lib
if len(sys.argv) > 1: _lib.LIB_PATH = os.path.abspath(sys.argv[1])
from hyperscanning_signal_analysis_b200 import mtmvar as mv
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()): return f(*a, **k)
out = {}
for name in ("mvar_cfg2_windows.npz", "mvar_cfg2_windows_lp40.npz"):
    g = np.load(os.path.join(G, name))
    for k in range(g["windows"].shape[0]):
        A, V = mv.ar_coeff(g["windows"][k], int(g["p"]))
        ff = quiet(mv.full_freq_dtf, g["windows"][k], g["freqs"], float(g["fs"]), optimal_model_order=int(g["p"]))
        out[f"{name[:-4]}[{k}]"] = dict(cond=float(g["cond"][k]), A=rel(A, g["A"][k]), V=rel(V, g["V"][k]), ffdtf=rel(ff, g["ffdtf"][k]))
g = np.load(os.path.join(G, "mvar_pcoh.npz"))
w = np.load(os.path.join(G, "mvar_cfg2_windows.npz"))["windows"][0]
dd = quiet(mv.direct_dtf, w, g["freqs"], 256.0, optimal_model_order=8)
out["ddtf_w0"] = rel(dd, g["w0_ddtf"])
c1 = np.load(os.path.join(G, "mvar_cfg1.npz"))
A, V = mv.ar_coeff(c1["x"], int(c1["p"])) if "x" in c1.files else (None, None)
if A is not None: out["cfg1"] = dict(A=rel(A, c1["A"]), V=rel(V, c1["V"]))
for k, v in out.items(): print(k, v)
