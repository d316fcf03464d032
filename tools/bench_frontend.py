"""Front-end throughput (cfg4-style): IIR filtfilt cascade, FIR decimation, multitaper PSD on the GPU, with the
SciPy / oracle CPU path timed next to it on a bounded sample.  Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import signal
from hyperscanning_signal_analysis_b200 import frontend, psd as gpsd, synth, _lib
from oracle import frontend_oracle as fo

seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 600
n_ch = 38
fs = 1024.0
x = synth.dyad_eeg(seed=7, m=n_ch, fs=fs, n_samples=int(seconds * fs), drift=True)
filt = fo.design_eeg_filters(fs, 1.0, 64.0)
flt3 = list(filt[:3])
xd = torch.from_numpy(x).cuda()
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts))

work = xd.clone()
def run_filt():
    work.copy_(xd)
    frontend.filtfilt_cascade_(work, flt3, remove_dc=True)
t_copy = timeit(lambda: work.copy_(xd))
t_filt = timeit(run_filt) - t_copy
y = frontend.decimate_dev(work, 8)
t_dec = timeit(lambda: frontend.decimate_dev(work, 8))
N = x.shape[1]
hbm = peaks.get("hbm_gbs", 6553.9)
# algorithmic bytes: mean pass (8 B) + 6 sweeps x (read 8 + write 8) per sample
bytes_filt = n_ch * N * (8 + 6 * 16)
bytes_dec = n_ch * (N * 8 + (N // 8) * 8)
out = {"signal": f"{n_ch} ch x {seconds} s @ {fs:.0f} Hz float64",
       "filtfilt_cascade_ms": t_filt, "filtfilt_Msamples_per_s": n_ch * N / t_filt * 1e-3,
       "filtfilt_algorithmic_GBs": bytes_filt / t_filt * 1e-6, "filtfilt_frac_of_measured_hbm": bytes_filt / t_filt * 1e-6 / hbm,
       "decimate_q8_ms": t_dec, "decimate_algorithmic_GBs": bytes_dec / t_dec * 1e-6, "decimate_frac_of_measured_hbm": bytes_dec / t_dec * 1e-6 / hbm}
# multitaper: 64-s segments at 128 Hz of the decimated signal
seg = 8192
nseg = y.shape[1] // seg
segs = y[:, : nseg * seg].reshape(n_ch * nseg, seg).contiguous()
gpsd.psd_multitaper_dev(segs, 128.0, 1.0, 30.0, 2.0); torch.cuda.synchronize()
t_psd = timeit(lambda: gpsd.psd_multitaper_dev(segs, 128.0, 1.0, 30.0, 2.0), reps=3)
K = gpsd._tapers(seg, 2.0 * seg / (2 * 128.0))[0].shape[0]
flops_psd = segs.shape[0] * ((K + 1) // 2) * 5.0 * seg * np.log2(seg)
out.update({"psd_segments": int(segs.shape[0]), "psd_tapers": int(K), "psd_ms": t_psd, "psd_segments_per_s": segs.shape[0] / t_psd * 1e3,
            "psd_fft_TFLOPs": flops_psd / t_psd * 1e-9})
# CPU on a bounded sample
xs = x[:4, : int(60 * fs)]
t0 = time.perf_counter(); ref = fo.apply_filters_iir(xs, filt); t_cpu = time.perf_counter() - t0
out["cpu_filtfilt_Msamples_per_s_1core"] = xs.size / t_cpu * 1e-6
t0 = time.perf_counter(); signal.decimate(ref, 8, ftype="fir", zero_phase=True, axis=1); out["cpu_decimate_Msamples_per_s_1core"] = xs.size / (time.perf_counter() - t0) * 1e-6
sn = segs[:2].cpu().numpy()
t0 = time.perf_counter(); fo.psd_multitaper(sn, 128.0, 1.0, 30.0, 2.0); out["cpu_psd_segments_per_s_1core"] = 2 / (time.perf_counter() - t0)
got = work[:4, : int(60 * fs)].cpu().numpy()
out["launches"] = _lib.launch_count()
print(json.dumps(out))
