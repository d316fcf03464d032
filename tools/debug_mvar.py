import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hyperscanning_signal_analysis_b200 import mtmvar as mv
from oracle import mvar_oracle as mo
def rel(a,b): return float(np.abs(a-b).max()/np.abs(b).max())
for (m,n,p) in ((40,700,10),(40,700,8),(38,700,10),(40,512,8),(38,512,9),(19,512,20),(39,640,3)):
    rng = np.random.default_rng(100 + m + p)
    x = rng.standard_normal((m, n)); x[:, 1:] += 0.5 * x[:, :-1]; x += 0.2 * rng.standard_normal((m, m)) @ x
    Rr = mo.lag_covariances(x, p)
    t, off, trials, mm, nn = mv._window_tensor(x)
    R = mv.batched_lagcov(t, off, n, 1, 1, m, n, p)
    eR = [rel(R[0,l].cpu().numpy(), Rr[l]) for l in range(p+1)]
    A, V, _, st = mv.batched_yw_solve(torch.from_numpy(Rr[None]).cuda())
    Ar, Vr = mo.ar_coeff(x, p)
    eA = [rel(A[0,:,:,k].cpu().numpy(), Ar[:,:,k]) for k in range(p)]
    print(m,n,p,'R err max %.1e'%max(eR), ['%.0e'%e for e in eR], 'A err', ['%.0e'%e for e in eA], 'V %.1e'%rel(V[0].cpu().numpy(),Vr), st.tolist())
