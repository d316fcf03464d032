"""Result exchange kernel (csrc/gather_kernels.cu) and the host-buffer plan, on one GPU.

The push kernel is exercised with the GPU's own second buffer standing in for a peer mapping (the store path is the same
instruction stream; the cross-GPU run is bench.py --gpus N and tools/gather_probe.py under torchrun)."""
import ctypes as C

import numpy as np
import pytest

from conftest import TOL_MODEL, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    from hyperscanning_signal_analysis_b200 import _lib
    return torch, _lib, _lib.load()


@pytest.mark.parametrize("count", [2, 510, 4096, 1_000_002])
def test_push_kernel_copies_to_every_peer(env, count):
    torch, _lib, lib = env
    src = torch.randn(count, dtype=torch.float64, device="cuda")
    peers = [torch.zeros(count + 2, dtype=torch.float64, device="cuda") for _ in range(3)]
    arr = (C.c_void_p * 3)(*[p.data_ptr() for p in peers])
    _lib.check(lib.hs_gather_push_f64(src.data_ptr(), count, None, arr, 3, 4, torch.cuda.current_stream().cuda_stream), "push")
    torch.cuda.synchronize()
    for p in peers:
        assert torch.equal(p[:count], src) and float(p[count:].abs().max()) == 0.0      # bit-exact, no overrun
    arr2 = (C.c_void_p * 3)(*[p.data_ptr() for p in peers])
    for p in peers:
        p.zero_()
    _lib.check(lib.hs_gather_push_ce(src.data_ptr(), count, arr2, 3, torch.cuda.current_stream().cuda_stream), "push_ce")
    torch.cuda.synchronize()
    for p in peers:
        assert torch.equal(p[:count], src)


def test_push_rejects_bad_arguments(env):
    torch, _lib, lib = env
    src = torch.zeros(8, dtype=torch.float64, device="cuda")
    arr = (C.c_void_p * 1)(src.data_ptr())
    assert lib.hs_gather_push_f64(src.data_ptr(), 3, None, arr, 1, 4, None) != 0          # odd count
    assert lib.hs_gather_push_f64(src.data_ptr() + 8, 2, None, arr, 1, 4, None) != 0      # unaligned source
    assert lib.hs_gather_push_f64(src.data_ptr(), 2, None, None, 0, 4, None) != 0         # no destination
    assert lib.hs_gather_push_f64(src.data_ptr(), 0, None, None, 0, 4, None) == 0         # empty push is a no-op


def test_sharded_ffdtf_single_rank_matches_batched_call(env):
    """ShardedFfdtf at world size 1 (no peers): chunked, SM-limited path == one hs_mvar_ffdtf_f64 call, bit for bit."""
    torch, _lib, lib = env
    from hyperscanning_signal_analysis_b200 import mtmvar, sharding, synth
    n_units, T, W, p = 3, 4096, 512, 8
    x = np.stack([synth.dyad_eeg(seed=40 + u, n_samples=T, line_amp=0.0) for u in range(n_units)])
    starts = np.linspace(0, T - W, 7, dtype=int)
    freqs = np.linspace(0, 128, 24, endpoint=False)
    xd = torch.from_numpy(x).cuda()
    sh = sharding.ShardedFfdtf(n_units, 38, T, W, starts, freqs, 256.0, p, units_per_chunk=2)
    sh.step(xd)
    sh.finish()
    torch.cuda.synchronize()
    got = sh.result.cpu().numpy()
    assert got.shape == (n_units * 7, 38, 38, 24)
    for u in range(n_units):
        ref = mtmvar.windowed_ffdtf(x[u], starts, W, freqs, 256.0, p).cpu().numpy()
        assert np.array_equal(got[u * 7:(u + 1) * 7], ref)
    sh.close()


def test_sm_limit_changes_partition_not_results(env):
    torch, _lib, lib = env
    from hyperscanning_signal_analysis_b200 import mtmvar, synth
    x = synth.dyad_eeg(seed=3, n_samples=4096, line_amp=0.0)
    starts = np.arange(0, 4096 - 512 + 1, 256)
    freqs = np.linspace(0, 128, 64, endpoint=False)
    a = mtmvar.windowed_ffdtf(x, starts, 512, freqs, 256.0, 8).cpu().numpy()
    try:
        _lib.check(lib.hs_set_compute_sm_limit(100), "limit")
        b = mtmvar.windowed_ffdtf(x, starts, 512, freqs, 256.0, 8).cpu().numpy()
    finally:
        lib.hs_set_compute_sm_limit(0)
    assert relerr(b, a) < 1e-13          # row sums are added in a different partition order: rounding only
    assert lib.hs_set_compute_sm_limit(-1) != 0


def test_plan_output_modes(env):
    """FfdtfPlan.run: page-locked out, pageable out, plan-owned out all deliver the same bits; bad `out` is rejected."""
    torch, _lib, lib = env
    from hyperscanning_signal_analysis_b200 import mtmvar, synth
    T, W, p, nf = 8192, 512, 8, 32
    x = synth.dyad_eeg(seed=9, n_samples=T, line_amp=0.0)
    starts = np.arange(0, T - W + 1, 128).astype(np.int64)          # 61 windows: several ramped chunks
    freqs = np.linspace(0, 128, nf, endpoint=False)
    ref = mtmvar.windowed_ffdtf(x, starts, W, freqs, 256.0, p).cpu().numpy()
    plan = mtmvar.FfdtfPlan(len(starts), 38, W, p, nf, T)
    pinned = torch.empty((len(starts), 38, 38, nf), dtype=torch.float64).pin_memory().numpy()
    assert plan.run(x, starts, freqs, 256.0, out=pinned) is pinned
    pageable = np.full((len(starts), 38, 38, nf), np.nan)
    plan.run(x, starts, freqs, 256.0, out=pageable)
    own = plan.run(x, starts, freqs, 256.0)
    assert own.shape == ref.shape and not own.flags["OWNDATA"]
    for got in (pinned, pageable, own):
        assert relerr(got, ref) < 1e-13
    assert np.array_equal(pinned, pageable) and np.array_equal(pinned, own)
    fewer = plan.run(x, starts[:5], freqs, 256.0)
    assert fewer.shape[0] == 5 and np.array_equal(fewer, pinned[:5])
    for bad in (np.empty((len(starts), 38, 38, nf), dtype=np.float32), np.empty((len(starts) - 1, 38, 38, nf)),
                np.empty((len(starts), 38, 38, 2 * nf))[..., ::2], [[0.0]]):
        with pytest.raises(ValueError):
            plan.run(x, starts, freqs, 256.0, out=bad)
    with pytest.raises(_lib.HsError):
        plan.run(x, np.array([T - 10]), freqs, 256.0)               # window outside the signal: C side refuses, plan stays usable
    again = plan.run(x, starts, freqs, 256.0)
    assert np.array_equal(again, pinned)
    plan.close()
    with pytest.raises(_lib.HsError):
        plan.run(x, starts, freqs, 256.0)
