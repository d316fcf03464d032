#!/usr/bin/env python
"""MVAR+DTF windows/sec (2x19 ch, p=8, 256 bins) -- BASELINE.json's metric.

One "step" = one pass of the hot path (lag covariances -> LWR Yule-Walker -> A(f)^-1 ->
|H|^2 -> ffDTF) over one 10-minute synthetic TALK task: 599 windows of 2 s at 50 % overlap,
38 channels, p = 8, 256 bins (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1: launched by torchrun, one rank per GPU; every rank owns its own dyad (weak scaling,
no data-path collective -- windows of different dyads are independent).  The final result
all-gather over NCCL that north_star describes is timed separately and reported under
"allgather" (it is not part of `value`; see DESIGN.md "Multi-GPU").
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M, WIN, HOP, P, F, FS = 38, 512, 256, 8, 256, 256.0
TASK_SECONDS = 600
# cfg3 (BASELINE.json configs[2]): 64 dyads x 3 tasks; BASELINE gives no task length, SURVEY 8d recommends 120 s
# (119 windows per task, 22 848 windows, 67.6 GB of float64 ffDTF: fits every GPU after the all-gather)
CFG3_DYADS, CFG3_TASKS, CFG3_SECONDS = 64, ("SECORE", "MOVIE", "TALK"), 120
CFG3_WINDOWS = (CFG3_SECONDS * 256 - WIN) // HOP + 1
# dram__bytes_read.sum + dram__bytes_write.sum of transfer_mma_kernel per 599-window launch, ncu --set full (profiles/)
K5_NCU_DRAM_BYTES_PER_599 = 72.581888e6 + 1718.427e6
METRIC = "MVAR+DTF windows/sec (2x19ch, p=8, 256 bins)"
UNIT = "windows/s"


# ------------------------------------------------------------------ workload
def make_task(seed, seconds=TASK_SECONDS):
    """Filtered synthetic task (m, T) + window starts + frequency grid."""
    from scipy import signal
    from hyperscanning_signal_analysis_b200 import synth
    x = synth.cfg2_raw(seed=seed, seconds=seconds)
    # notch 50 Hz (Q=30) + band 1-64 Hz, the reference's production values
    # (scripts/export_dyade_to_ncdf_by_task_batch.py:27-31); setup only, outside every timed region
    y = x - x.mean(axis=1, keepdims=True)
    for b, a in (signal.iirnotch(50.0, 30.0, fs=FS), signal.butter(2, 64.0, "low", fs=FS), signal.butter(2, 1.0, "high", fs=FS)):
        y = signal.filtfilt(b, a, y, axis=1)
    T = y.shape[1]
    n_win = (T - WIN) // HOP + 1
    starts = np.linspace(0, T - WIN, n_win, dtype=int).astype(np.int64)      # _create_windows, eeg_alpha_ibi_ffdtf.py:513
    freqs = np.linspace(0.0, FS / 2, F, endpoint=False)
    return np.ascontiguousarray(y), starts, freqs


def flops_per_window():
    """Algorithmic FLOPs per window, SURVEY.md 8(d) (reference's algorithm counts)."""
    m, n, p, f = M, WIN, P, F
    k3 = 2 * m * m * sum(n - L for L in range(p + 1))
    mp = m * p
    k4 = (2 / 3) * mp ** 3 + 2 * mp ** 2 * m + 2 * m * m * mp
    k5_asm = 4 * p * m * m * f
    k5_inv = 8 * m ** 3 * f
    k5_red = 5 * m * m * f
    return {"lagcov": k3, "yule_walker": k4, "transfer": k5_asm + k5_inv, "normalise": k5_red,
            "total": k3 + k4 + k5_asm + k5_inv + k5_red}


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            for nm, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        pw = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.samples), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_baseline(y, starts, freqs, budget_s=20.0):
    """Oracle port of full_freq_dtf (same per-bin np.linalg.inv loop as the reference) on every host core:
    one worker process per core (1 BLAS thread each), windows dealt evenly.  Bounded sample of the same task."""
    import tempfile
    cores = os.cpu_count() or 1
    from oracle import mvar_oracle as mo
    t0 = time.perf_counter()
    mo.full_freq_dtf(y[:, :WIN], freqs, FS, optimal_model_order=P)
    t_one = time.perf_counter() - t0
    n = int(min(len(starts), max(cores, budget_s * cores / max(t_one, 1e-3))))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "windows.npz")
        np.savez(path, windows=np.stack([y[:, s:s + WIN] for s in starts[:n]]), freqs=freqs)
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
        bounds = np.linspace(0, n, cores + 1).astype(int)
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "oracle", "cpu_worker.py"), path, str(bounds[i]), str(bounds[i + 1]),
                                   str(FS), str(P)], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, env=env)
                 for i in range(cores)]
        for pr in procs:
            assert pr.stdout.readline().strip() == "ready"
        t0 = time.perf_counter()
        for pr in procs:
            pr.stdin.write("go\n")
            pr.stdin.flush()
        for pr in procs:
            assert pr.stdout.readline().startswith("done")
        dt = time.perf_counter() - t0
        for pr in procs:
            pr.wait()
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} of {len(starts)} windows of the same task, oracle/mvar_oracle.full_freq_dtf (NumPy port of src/mtmvar.py), "
                      f"{cores} processes x 1 BLAS thread; single-process {1.0 / t_one:.2f} windows/s",
            "serial_value": 1.0 / t_one,       # one process, one window at a time: how the reference's run_pipeline loop runs
            "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    y, starts, freqs = make_task(20260101)
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(y, starts, freqs, budget_s=8.0)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    n = int(vals[-1]["sample"].split()[0])
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean([r["seconds"] for r in vals])), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_config(args.gpus), "windows_per_step": n,
                       "sample": "each step = a bounded sample of these windows on the host CPU (rank 0 only)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": vals[-1]["cores"], "kind": "port", "sample": vals[-1]["sample"],
                             "serial_value": vals[-1]["serial_value"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------ B200 arm
def workload_config(world):
    """config.workload of BOTH arms (the reference arm runs a bounded sample of the same windows on the host)."""
    if world == 1:
        return ("cfg2: sliding-window ffDTF, 599 windows x (38 ch x 512 samples) of one 600 s task per GPU, p=8, F=256, hop 256")
    return (f"cfg3: 64 dyads x 3 tasks x {CFG3_SECONDS} s ({CFG3_WINDOWS} windows each of 38 ch x 512, p=8, F=256, hop 256) sharded by dyad over "
            f"{world} GPUs, full float64 result all-gathered over NVLink inside the timed step")


def measure_cfg2(args, lib, mtmvar, _lib, torch, dist, world, rank, local, with_stages=True):
    """cfg2 on this rank's GPU: device-resident throughput, per-stage times + roofline of the dominant kernel, e2e."""
    import ctypes as C
    y, starts, freqs = make_task(20260101 + 3 * rank)
    n_win = len(starts)
    T = y.shape[1]
    x_d = torch.from_numpy(y).cuda()
    st_d = torch.from_numpy(starts).cuda()
    fr_d = torch.from_numpy(freqs).cuda()
    out_d = torch.empty((n_win, M, M, F), dtype=torch.float64, device="cuda")
    status = torch.zeros(n_win, dtype=torch.int32, device="cuda")
    ws = torch.empty(lib.hs_mvar_ffdtf_ws_bytes(n_win, M, P, F), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def step():
        flush.zero_()                                        # L2 flush (256 MiB > 126 MB L2), inside the timed region
        _lib.check(lib.hs_mvar_ffdtf_f64(x_d.data_ptr(), st_d.data_ptr(), T, n_win, M, WIN, P, fr_d.data_ptr(), F, FS,
                                         out_d.data_ptr(), None, None, status.data_ptr(), ws.data_ptr(), stream.cuda_stream),
                   "hs_mvar_ffdtf_f64")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        launches = _lib.launch_count() - launches0
        # nvidia-smi answers in ~50-100 ms, the K timed steps take ~8 ms each: keep the identical step loop running
        # (untimed) so that the clock / throttle samples describe this load and not an idle GPU
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end:
            for _ in range(4):
                step()
            torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    assert int(status.max().item()) == 0, "singular window in the synthetic workload"
    res = {"value": world * n_win * args.steps / (total_ms * 1e-3), "total_ms": total_ms, "launches": int(launches), "n_win": n_win,
           "clocks": dict(clk.summary(), window="timed steps + 1.5 s of the same step loop continued untimed")}

    # ---------------- per-stage device times (same stream, CUDA events), K5 = dominant kernel
    fl = flops_per_window()
    if with_stages:
        R = torch.empty((n_win, P + 1, M, M), dtype=torch.float64, device="cuda")
        A = torch.empty((n_win, M, M, P), dtype=torch.float64, device="cuda")
        V = torch.empty((n_win, M, M), dtype=torch.float64, device="cuda")
        yw_ws = torch.empty(lib.hs_yw_ws_bytes(n_win, M, P), dtype=torch.uint8, device="cuda")
        tr_ws = torch.empty(lib.hs_transfer_ws_bytes(n_win, M, P, F), dtype=torch.uint8, device="cuda")
        sp = stream.cuda_stream
        stages = {
            "lagcov": lambda: lib.hs_lagcov_f64(x_d.data_ptr(), st_d.data_ptr(), T, n_win, 1, M, WIN, P, R.data_ptr(), sp),
            "yule_walker": lambda: lib.hs_yw_solve_f64(R.data_ptr(), n_win, M, P, A.data_ptr(), V.data_ptr(), None, status.data_ptr(), yw_ws.data_ptr(), sp),
            "transfer": lambda: lib.hs_transfer_dtf_f64(A.data_ptr(), fr_d.data_ptr(), F, FS, n_win, M, P, None, None, out_d.data_ptr(), None,
                                                        status.data_ptr(), tr_ws.data_ptr(), sp),
            "transfer+normalise": lambda: lib.hs_transfer_dtf_f64(A.data_ptr(), fr_d.data_ptr(), F, FS, n_win, M, P, None, None, None, out_d.data_ptr(),
                                                                  status.data_ptr(), tr_ws.data_ptr(), sp),
        }
        stage_ms = {}
        k5_kernel_ms = []
        for name, fn in stages.items():
            ts = []
            lib.hs_timing_enable(1 if name == "transfer+normalise" else 0)
            for i in range(3 + max(3, args.steps)):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _lib.check(fn(), name)
                b.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(a.elapsed_time(b))
                    if name == "transfer+normalise":
                        kms = C.c_double(0.0)
                        _lib.check(lib.hs_timing_last_k5_ms(C.byref(kms)), "hs_timing_last_k5_ms")
                        k5_kernel_ms.append(kms.value)
            stage_ms[name] = float(np.mean(ts))
        lib.hs_timing_enable(0)
        # the dominant kernel alone (CUDA events recorded by the library around its launch, same stream)
        k5_ms = float(np.mean(k5_kernel_ms))
        stage_ms["transfer_kernel"] = k5_ms
        off = lib.hs_transfer_ws_flag_offset(n_win, M, P, F)
        flagged = int(tr_ws[off:off + 4].view(torch.int32).item())
        tf = C.c_double(0.0)
        _lib.check(lib.hs_measure_dfma_tflops(C.byref(tf), flush.data_ptr(), 5), "hs_measure_dfma_tflops")
        dfma_peak = tf.value
        k5_tflops = fl["transfer"] * n_win / (k5_ms * 1e-3) * 1e-12
        peaks = load_peaks()
        res["roofline"] = {
            "kernel": "transfer_mma_kernel: A(f) assembly + complex 38x38 block Gauss-Jordan on the FP64 tensor pipe (DMMA m8n8k4) + |H|^2, "
                      "599*256 matrices per launch",
            "bound": "tensor", "pipe": "fp64 (DMMA m8n8k4: the only tensor-pipe type for float64; tcgen05 has no f64 kind)",
            "achieved": k5_tflops, "peak": dfma_peak, "unit": "TFLOP/s", "frac": k5_tflops / dfma_peak if dfma_peak else None,
            "peak_source": "hs_measure_dfma_tflops: register-only DFMA loop on all SMs, measured in this run; DMMA m8n8k4 peaks at the same "
                           "rate (tools/fp64_peak.cu). MEASURED_PEAKS.json has no FP64 figure (its hbm_gbs is %s)" % peaks.get("hbm_gbs"),
            "flops_per_launch": fl["transfer"] * n_win, "ms_per_launch": k5_ms,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the ncu --set full capture under profiles/, scaled to
            # this launch's window count; algorithmic bytes: A (m*m*p*8) in, |H|^2 (m*m*F*8) out per window
            "traffic": K5_NCU_DRAM_BYTES_PER_599 * n_win / 599.0,
            "algorithmic_bytes_per_launch": n_win * (M * M * P * 8 + M * M * F * 8),
            "algorithmic_flops_per_matrix": "4*p*m^2 + 8*m^3 (SURVEY 8d: assembly + complex LU/inverse as the reference computes it)",
            "matrices_redone_with_pivoting": flagged,
            "stages_ms": stage_ms,
            "stages_tflops": {"lagcov": fl["lagcov"] * n_win / (stage_ms["lagcov"] * 1e-3) * 1e-12,
                              "yule_walker": fl["yule_walker"] * n_win / (stage_ms["yule_walker"] * 1e-3) * 1e-12,
                              "transfer_kernel": k5_tflops},
            "finalize_hbm_gbs": 2 * n_win * M * M * F * 8 / (max(stage_ms["transfer+normalise"] - k5_ms, 1e-6) * 1e-3) * 1e-9,
            "step_tflops": fl["total"] * n_win * args.steps * world / (total_ms * 1e-3) * 1e-12}
        del R, A, V, yw_ws, tr_ws

    # ---------------- e2e through the host-buffer API (pinned NumPy in, pinned NumPy out)
    x_pin = torch.from_numpy(y).pin_memory()
    out_pin = torch.empty((n_win, M, M, F), dtype=torch.float64).pin_memory()
    plan = mtmvar.FfdtfPlan(n_win, M, WIN, P, F, T)
    x_np, out_np = x_pin.numpy(), out_pin.numpy()
    plan.run(x_np, starts, freqs, FS, out=out_np)
    e2e_steps = max(2, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.run(x_np, starts, freqs, FS, out=out_np)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": world * n_win * e2e_steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(y.nbytes + starts.nbytes + freqs.nbytes),
           "d2h_bytes_per_step": int(out_np.nbytes + 4 * n_win), "steps": e2e_steps,
           "api": "mtmvar.FfdtfPlan.run -> hs_plan_mvar_ffdtf_host (chunked H2D/compute/D2H on 3 streams), host wall clock incl. final sync; "
                  "one cfg2 task (599 windows) per GPU",
           "host_numa_binding": res.get("numa", "none")}
    assert abs(float(out_np[0].sum()) - M) < 1e-6
    # the same call with the result left in the plan's own pinned buffer (what run(out=None) returns a view of)
    view = plan.run(x_np, starts, freqs, FS)              # first call page-locks the plan's own result buffer (not timed)
    t0 = time.perf_counter()
    view = plan.run(x_np, starts, freqs, FS)
    e2e["value_plan_owned_output"] = n_win / (time.perf_counter() - t0)
    assert abs(float(view[0].sum()) - M) < 1e-6
    del view
    plan.close()
    # PCIe ceiling of this box for the e2e number: one plain device -> pinned-host copy of the result (same buffers)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out_pin.copy_(out_d, non_blocking=True)
    torch.cuda.synchronize()
    a.record()
    out_pin.copy_(out_d, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    d2h_ms = a.elapsed_time(b)
    e2e["pcie_d2h_gbs_measured"] = out_np.nbytes / (d2h_ms * 1e-3) * 1e-9
    e2e["pcie_bound_windows_per_s"] = world * n_win / (d2h_ms * 1e-3)
    e2e["frac_of_pcie_bound"] = e2e["value"] / e2e["pcie_bound_windows_per_s"]
    res["e2e"] = e2e
    res["task"] = (y, starts, freqs)
    return res


def _time_dev(torch, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


def measure_frontend(args, lib, _lib, torch):
    """cfg4 (BASELINE.json configs[3]): 38 ch x 1 h @ 1024 Hz: DC removal + notch/low/high-pass filtfilt (K1), decimate q = 8
    (K2), multitaper PSD of every 64-s segment (K6).  Device-resident times with CUDA events; the 1.12 GB signal is larger
    than L2.  CPU legs: SciPy / the oracle on a bounded sample, one core."""
    from scipy import signal
    from hyperscanning_signal_analysis_b200 import frontend, psd as gpsd, synth
    n_ch, fs, seconds = 38, 1024.0, 3600
    N = int(seconds * fs)
    # the generator is a sequential VAR simulation (~100 s for an hour): a 10-min block tiled 6 times with fresh noise is the
    # same workload for the filters (setup only)
    base = synth.dyad_eeg(seed=7, m=n_ch, fs=fs, n_samples=N // 6, drift=True)
    rng = np.random.default_rng(11)
    x = np.concatenate([base + 0.5 * rng.standard_normal(base.shape) for _ in range(6)], axis=1)
    # production filter values (scripts/export_dyade_to_ncdf_by_task_batch.py:27-31), designed as dataloader._design_eeg_filters does
    flt3 = [signal.iirnotch(50.0, 30.0, fs=fs), signal.butter(2, 64.0, "low", fs=fs), signal.butter(2, 1.0, "high", fs=fs)]
    xd = torch.from_numpy(x).cuda()
    work = torch.empty_like(xd)
    hbm = load_peaks().get("hbm_gbs", 6553.9)

    def run_filt():
        work.copy_(xd)
        frontend.filtfilt_cascade_(work, flt3, remove_dc=True)

    t_copy = _time_dev(torch, lambda: work.copy_(xd), reps=3)
    t_filt = _time_dev(torch, run_filt, reps=3) - t_copy
    y = frontend.decimate_dev(work, 8)
    t_dec = _time_dev(torch, lambda: frontend.decimate_dev(work, 8), reps=3)
    bytes_filt = n_ch * N * (8 + 6 * 16)           # mean pass + 6 sweeps x (read + write): SURVEY 8d's unfused count
    bytes_dec = n_ch * (N * 8 + (N // 8) * 8)
    out = {"workload": f"cfg4: {n_ch} ch x {seconds} s @ {fs:.0f} Hz float64 ({x.nbytes / 1e9:.2f} GB), notch/low/high filtfilt + decimate q=8 + "
                       "multitaper PSD per 64-s segment (bandwidth 2 Hz, 126 tapers, 1-30 Hz)",
           "filtfilt": {"ms": t_filt, "Msamples_per_s": n_ch * N / t_filt * 1e-3, "algorithmic_bytes": bytes_filt,
                        "roofline": {"bound": "hbm", "achieved": bytes_filt / t_filt * 1e-6, "peak": hbm, "unit": "GB/s",
                                     "frac": bytes_filt / t_filt * 1e-6 / hbm}},
           "decimate_q8": {"ms": t_dec, "algorithmic_bytes": bytes_dec,
                           "roofline": {"bound": "hbm", "achieved": bytes_dec / t_dec * 1e-6, "peak": hbm, "unit": "GB/s",
                                        "frac": bytes_dec / t_dec * 1e-6 / hbm}}}
    seg = 8192
    nseg = y.shape[1] // seg
    segs = y[:, : nseg * seg].reshape(n_ch * nseg, seg).contiguous()
    t_psd = _time_dev(torch, lambda: gpsd.psd_multitaper_dev(segs, 128.0, 1.0, 30.0, 2.0), reps=3)
    K = gpsd._tapers(seg, 2.0 * seg / (2 * 128.0))[0].shape[0]
    flops_psd = segs.shape[0] * ((K + 1) // 2) * 5.0 * seg * np.log2(seg)
    out["multitaper_psd"] = {"segments": int(segs.shape[0]), "tapers": int(K), "n": seg, "ms": t_psd, "segments_per_s": segs.shape[0] / t_psd * 1e3,
                             "fft_flops": flops_psd, "fft_tflops": flops_psd / t_psd * 1e-9}
    # odd lengths (real movie segments are not powers of two, io_utils.py:131): 60 s and a prime length
    for n_odd in (7680, 8191):
        so = y[:, : n_odd * (y.shape[1] // n_odd)].reshape(-1, n_odd)[:342].contiguous()
        t_o = _time_dev(torch, lambda: gpsd.psd_multitaper_dev(so, 128.0, 1.0, 30.0, 2.0), reps=2, warm=1)
        out["multitaper_psd"][f"ms_n{n_odd}_x{so.shape[0]}"] = t_o
    # CPU on a bounded sample (one core): the only place this function touches oracle/
    from oracle import frontend_oracle as fo
    xs = x[:4, : int(120 * fs)]
    filt = fo.design_eeg_filters(fs, 1.0, 64.0)
    t0 = time.perf_counter()
    ref = fo.apply_filters_iir(xs, filt)
    t_cpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    signal.decimate(ref, 8, ftype="fir", zero_phase=True, axis=1)
    t_cpu_dec = time.perf_counter() - t0
    sn = segs[:2].cpu().numpy()
    t0 = time.perf_counter()
    fo.psd_multitaper(sn, 128.0, 1.0, 30.0, 2.0)
    t_cpu_psd = time.perf_counter() - t0
    out["cpu_baseline"] = {"kind": "port", "cores": 1, "sample": "4 ch x 120 s (SciPy filtfilt cascade + decimate), 2 segments (oracle multitaper)",
                           "filtfilt_Msamples_per_s": xs.size / t_cpu * 1e-6, "decimate_Msamples_per_s": xs.size / t_cpu_dec * 1e-6,
                           "psd_segments_per_s": 2 / t_cpu_psd}
    out["speedup_vs_1core"] = {"filtfilt": out["filtfilt"]["Msamples_per_s"] / out["cpu_baseline"]["filtfilt_Msamples_per_s"],
                               "psd": out["multitaper_psd"]["segments_per_s"] / out["cpu_baseline"]["psd_segments_per_s"]}
    return out


def measure_cfg5(args, lib, _lib, torch):
    """cfg5 (BASELINE.json configs[4]): 2 x 64 ch, p = 15, 512 bins, covariances averaged over 100 epochs per window -- the
    dense lag-covariance contraction (26.5 GFLOP per window) plus the m = 128 solve and inverses (generic path)."""
    import ctypes as C
    from hyperscanning_signal_analysis_b200 import mtmvar, synth
    m, n, trials, p, nf, n_windows = 128, 512, 100, 15, 512, 8
    ep = synth.cfg5_epochs(n_windows=n_windows)                                   # (w, m, n, trials)
    x = torch.from_numpy(np.ascontiguousarray(ep.transpose(0, 3, 1, 2))).cuda()   # (w, trials, m, n)
    offs = (torch.arange(n_windows * trials, dtype=torch.int64, device="cuda") * (m * n)).contiguous()
    freqs = np.linspace(0, 128, nf, endpoint=False)
    fr = torch.from_numpy(freqs).cuda()
    sp = torch.cuda.current_stream().cuda_stream
    R = torch.empty((n_windows, p + 1, m, m), dtype=torch.float64, device="cuda")
    A = torch.empty((n_windows, m, m, p), dtype=torch.float64, device="cuda")
    V = torch.empty((n_windows, m, m), dtype=torch.float64, device="cuda")
    ff = torch.empty((n_windows, m, m, nf), dtype=torch.float64, device="cuda")
    status = torch.zeros(n_windows, dtype=torch.int32, device="cuda")
    yw_ws = torch.empty(lib.hs_yw_ws_bytes(n_windows, m, p), dtype=torch.uint8, device="cuda")
    tr_ws = torch.empty(lib.hs_transfer_ws_bytes(n_windows, m, p, nf), dtype=torch.uint8, device="cuda")
    t_k3 = _time_dev(torch, lambda: _lib.check(lib.hs_lagcov_f64(x.data_ptr(), offs.data_ptr(), n, n_windows, trials, m, n, p, R.data_ptr(), sp), "k3"), reps=3, warm=1)
    t_k4 = _time_dev(torch, lambda: _lib.check(lib.hs_yw_solve_f64(R.data_ptr(), n_windows, m, p, A.data_ptr(), V.data_ptr(), None, status.data_ptr(),
                                                                  yw_ws.data_ptr(), sp), "k4"), reps=3, warm=1)
    t_k5 = _time_dev(torch, lambda: _lib.check(lib.hs_transfer_dtf_f64(A.data_ptr(), fr.data_ptr(), nf, 256.0, n_windows, m, p, None, None, None, ff.data_ptr(),
                                                                      status.data_ptr(), tr_ws.data_ptr(), sp), "k5"), reps=3, warm=1)
    assert int(status.max().item()) == 0
    rows = ff.sum(dim=(2, 3))
    assert float((rows - 1).abs().max().item()) < 1e-9
    f_k3 = 2.0 * m * m * trials * sum(n - L for L in range(p + 1))
    mp = m * p
    f_k4 = (2 / 3) * mp ** 3 + 2 * mp ** 2 * m + 2 * m * m * mp
    f_k5 = (4 * p * m * m + 8 * m ** 3) * nf
    tf = C.c_double(0.0)
    _lib.check(lib.hs_measure_dfma_tflops(C.byref(tf), tr_ws.data_ptr(), 3), "hs_measure_dfma_tflops")
    peak = tf.value
    tot_ms = t_k3 + t_k4 + t_k5
    out = {"workload": f"cfg5: {n_windows} windows x (128 ch x 512 samples x 100 epochs), p=15, F=512",
           "windows_per_s": n_windows / tot_ms * 1e3, "ms": {"lagcov": t_k3, "yule_walker": t_k4, "transfer+normalise": t_k5},
           "lagcov_roofline": {"bound": "tensor", "pipe": "fp64 (DMMA m8n8k4)", "achieved": f_k3 * n_windows / t_k3 * 1e-9, "peak": peak, "unit": "TFLOP/s",
                               "frac": f_k3 * n_windows / t_k3 * 1e-9 / peak, "flops_per_window": f_k3},
           "yule_walker_tflops": f_k4 * n_windows / t_k4 * 1e-9, "transfer_tflops": f_k5 * n_windows / t_k5 * 1e-9,
           "transfer_frac_of_fp64_peak": f_k5 * n_windows / t_k5 * 1e-9 / peak}
    if not args.no_cpu:
        from oracle import mvar_oracle as mo
        t0 = time.perf_counter()
        mo.ar_coeff(ep[0], p)
        t_fit = time.perf_counter() - t0
        out["cpu_baseline"] = {"kind": "port", "cores": os.cpu_count(), "sample": "ar_coeff of 1 window (NumPy/BLAS threads as configured)",
                               "ar_coeff_windows_per_s": 1.0 / t_fit, "gpu_fit_windows_per_s": n_windows / (t_k3 + t_k4) * 1e3}
    return out


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def measure_cfg3(args, lib, _lib, torch, dist, world, rank, local):
    """cfg3: 64 dyads x 3 tasks sharded by dyad (sharding.shard_units), the full result all-gathered over NVLink chunk by
    chunk on a second stream while the next chunk computes (sharding.ShardedFfdtf).  Timed region = K x (all chunks of
    this rank + all pushes + rank barrier), CUDA events, max over ranks."""
    from scipy import signal
    from hyperscanning_signal_analysis_b200 import sharding, synth
    units = sharding.unit_table(CFG3_DYADS, CFG3_TASKS)                       # (dyad, task) in run_pipeline's order
    mine = sharding.shard_units(len(units), rank, world)
    assert len(units) % world == 0
    T = int(CFG3_SECONDS * FS)
    filt = (signal.iirnotch(50.0, 30.0, fs=FS), signal.butter(2, 64.0, "low", fs=FS), signal.butter(2, 1.0, "high", fs=FS))
    x_host = np.empty((len(mine), M, T), dtype=np.float64)
    for k, u in enumerate(mine):
        dyad, task = units[u]
        x = synth.cfg2_raw(seed=synth.cfg3_seed(dyad, CFG3_TASKS.index(task)), seconds=CFG3_SECONDS)
        yk = x - x.mean(axis=1, keepdims=True)
        for b, a in filt:
            yk = signal.filtfilt(b, a, yk, axis=1)
        x_host[k] = yk
    x_all = torch.from_numpy(x_host).cuda()
    starts = np.linspace(0, T - WIN, CFG3_WINDOWS, dtype=int).astype(np.int64)
    freqs = np.linspace(0.0, FS / 2, F, endpoint=False)
    push = args.gather
    out = {}

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(sh, steps, warm):
        for _ in range(warm):
            sh.step(x_all)
            sh.finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            sh.step(x_all)
            sh.finish()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), _lib.launch_count() - l0

    sh = sharding.ShardedFfdtf(len(mine), M, T, WIN, starts, freqs, FS, P, units_per_chunk=args.chunk_units, push=push,
                               push_ctas=args.push_ctas, buffer_mode=args.gather_buffer)
    if push == "multicast" and not sh.buf.multicast_ptr:
        raise SystemExit("bench.py: --gather multicast needs an NVSwitch multicast mapping (symmetric memory)")
    with ClockSampler(local) as clk:
        total_ms, launches = timed(sh, args.steps, max(args.warmup, 3))
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:                 # keep the same load running for the clock samples
            sh.step(x_all)
            sh.finish()
            torch.cuda.synchronize()
    assert int(sh.status.max().item()) == 0, "singular window in the synthetic workload"
    total_windows = sh.total_windows
    out["value"] = total_windows * args.steps / (total_ms * 1e-3)
    out["total_ms"] = total_ms
    out["launches"] = int(launches)
    out["clocks"] = dict(clk.summary(), window="timed steps + 1 s of the same step loop continued untimed")
    out["total_windows"] = int(total_windows)
    # every slot must hold ffDTF (rows sum to 1) and the same bits on every rank
    res = sh.result
    probe_w = torch.tensor([r * sh.win_local + k for r in range(world) for k in (0, sh.win_local // 2, sh.win_local - 1)], device="cuda")
    rows = res[probe_w].sum(dim=(2, 3))
    assert float((rows - 1.0).abs().max().item()) < 1e-9, "gathered slots do not hold ffDTF rows"
    chk = res[probe_w].double().sum(dim=(1, 2, 3)).contiguous()
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert bool(torch.equal(lo, hi)), "ranks disagree on the gathered result"
    recv = (world - 1) * sh.win_local * sh.per_win * 8
    ag = {"mode": push, "buffer": sh.buf.mode, "multicast_available": bool(sh.buf.multicast_ptr), "push_ctas": sh.push_ctas if push in ("p2p", "multicast") else 0,
          "compute_sm_limit": sh.sm_limit, "chunk_windows": args.chunk_units * CFG3_WINDOWS,
          "bytes_received_per_rank_per_step": int(recv), "gathered_bytes_per_gpu": int(total_windows * sh.per_win * 8),
          "what_is_gathered": "the full (windows, 38, 38, 256) float64 ffDTF of every (dyad, task, window): what the reference persists "
                              "(eeg_alpha_ibi_ffdtf.py:647-656)",
          "nvlink_in_gbs_per_rank": recv * args.steps / (total_ms * 1e-3) * 1e-9, "nvlink_peer_copy_gbs_reference": 770.0}
    # the same shard without any exchange, with the exchange as ONE NCCL all-gather after the compute (the baseline), and with
    # the other push mechanisms (each a few steps; `value` above is the --gather mode)
    extra_steps = max(2, min(args.steps, 3))
    others = [("none", "value_excl_allgather"), ("nccl", "value_nccl_allgather_after_compute"), ("ce", "value_push_copy_engines"),
              ("p2p", "value_push_store_kernel")]
    if sh.buf.multicast_ptr:
        others.append(("multicast", "value_push_multicast_kernel"))
    for mode, key in others:
        if mode == push:
            ag[key] = out["value"]
            continue
        sh.push = mode
        sh.sm_limit = sh.sm_total - sh.push_ctas if mode in ("p2p", "multicast") else 0
        _lib.check(lib.hs_set_compute_sm_limit(sh.sm_limit), "hs_set_compute_sm_limit")
        ms2, _ = timed(sh, extra_steps, 1)
        ag[key] = total_windows * extra_steps / (ms2 * 1e-3)
    sh.push = push
    # what bounds the step: this rank's compute (value_excl_allgather) or the bytes it must RECEIVE over NVLink
    link = 770.0
    t_gather = recv / (link * 1e9)
    t_compute = total_windows / ag["value_excl_allgather"]
    ag["bound"] = {"compute_s": t_compute, "nvlink_receive_s_at_770GBs": t_gather, "step_s": total_ms * 1e-3 / args.steps,
                   "frac_of_max(compute, receive)": max(t_compute, t_gather) / (total_ms * 1e-3 / args.steps)}
    out["allgather"] = ag
    sh.close()
    del sh, x_all
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from hyperscanning_signal_analysis_b200 import _lib, mtmvar

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    numa_cpus = None
    if world > 1 and os.environ.get("HS_BENCH_NUMA", "1") != "0":
        from hyperscanning_signal_analysis_b200 import sharding
        numa_cpus = sharding.bind_to_gpu_numa(local)      # pinned e2e buffers of every rank on its GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()

    if world == 1:
        r = measure_cfg2(args, lib, mtmvar, _lib, torch, dist, world, rank, local)
        value, total_ms, launches, clocks = r["value"], r["total_ms"], r["launches"], r["clocks"]
        config = {"workload": workload_config(1), "windows_per_step_per_gpu": r["n_win"], "parallelism": "single GPU",
                  "l2": "explicit 256 MiB memset before every step, inside the timed region; each step also writes 1.77 GB (> 126 MB L2)"}
        extra = {}
        if not args.no_extra:
            extra["frontend_cfg4"] = measure_frontend(args, lib, _lib, torch)
            extra["stress_cfg5"] = measure_cfg5(args, lib, _lib, torch)
    else:
        c3 = measure_cfg3(args, lib, _lib, torch, dist, world, rank, local)
        r = measure_cfg2(args, lib, mtmvar, _lib, torch, dist, world, rank, local)
        value, total_ms, launches, clocks = c3["value"], c3["total_ms"], c3["launches"], c3["clocks"]
        config = {"workload": workload_config(world), "windows_per_step": c3["total_windows"],
                  "parallelism": f"dyads sharded over {world} GPUs (shard_units), chunks of {args.chunk_units} tasks; result pushed to every peer "
                                 f"({args.gather}) on a second stream under the next chunk's compute; one rank barrier per step",
                  "l2": "no explicit flush: every step streams > 8 GB of results per GPU through the 126 MB L2"}
        extra = {"allgather": c3["allgather"],
                 "per_gpu_cfg2_no_collective": {"value": r["value"], "ms_per_step": r["total_ms"] / args.steps,
                                                "note": "one cfg2 task per GPU, no exchange (round 1's headline); for reference only"}}
        # scaling is strong: the cfg3 job (22 848 windows) is the same at every N > 1
    e2e = r["e2e"]
    e2e["host_numa_binding"] = ("rank bound to %d GPU-local CPUs" % len(numa_cpus)) if numa_cpus else "none"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        y, starts, freqs = r["task"]
        cpu = cpu_baseline(y, starts, freqs, budget_s=20.0)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": r["roofline"]}
        line.update(extra)
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="N=1: skip the cfg4 front-end and cfg5 stress legs")
    ap.add_argument("--gather", default="ce", choices=["ce", "p2p", "multicast", "nccl"],
                    help="N>1: how a finished chunk reaches the peers (own store kernel over peer mappings / NVSwitch multicast, copy engines, "
                         "or one NCCL all-gather after the compute)")
    ap.add_argument("--gather-buffer", default="auto", choices=["auto", "symm", "ipc"])
    ap.add_argument("--push-ctas", type=int, default=16)
    ap.add_argument("--chunk-units", type=int, default=0,
                    help="N>1: tasks per exchange chunk; 0 = automatic: 5 (x 119 windows ~ one cfg2 batch), except 2 around N = 4, where the NVLink "
                         "receive time is as long as the compute and a finer pipeline hides more of it (measured: N = 4 294 -> 315 k windows/s; "
                         "at N = 8, receive-bound, small chunks only add copies: 281 -> 229 k)")
    args = ap.parse_args()
    if args.chunk_units <= 0:
        args.chunk_units = 2 if 3 <= args.gpus <= 5 else 5
    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 on their own (NCCL prints its version
    # banner there when NCCL_DEBUG=VERSION is set in the environment) are sent to stderr for the duration of the run.
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_out, 1)
        os.close(real_out)
    lines = [ln for ln in buf.getvalue().splitlines() if ln.strip()]
    for ln in lines[:-1]:
        print(ln, file=sys.stderr)
    if lines:
        print(lines[-1])
        sys.stdout.flush()


if __name__ == "__main__":
    main()
