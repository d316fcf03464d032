"""CPU-baseline worker (TEST/BENCH INFRASTRUCTURE): runs the oracle port of full_freq_dtf over a
slice of windows.  Protocol: load + warm up, print 'ready', wait for a line on stdin, run, print
'done <seconds> <checksum>'.  Started by bench.py's cpu_baseline / --impl reference legs only."""
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

import numpy as np  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mvar_oracle as mo  # noqa: E402


def main():
    path, lo, hi, fs, p = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
    z = np.load(path)
    wins, freqs = z["windows"][lo:hi], z["freqs"]
    if len(wins):
        mo.full_freq_dtf(wins[0], freqs, fs, optimal_model_order=p)        # warm-up
    print("ready", flush=True)
    sys.stdin.readline()
    t0 = time.perf_counter()
    acc = 0.0
    for w in wins:
        acc += float(mo.full_freq_dtf(w, freqs, fs, optimal_model_order=p)[0, 0, 0])
    print(f"done {time.perf_counter() - t0:.6f} {acc:.6e}", flush=True)


if __name__ == "__main__":
    main()
