"""Multi-GPU layer: one process per GPU (torch.distributed), units sharded by dyad/task, no collective inside
the computation (every (dyad, task, window) is independent: run_pipeline's outer loops,
eeg_alpha_ibi_ffdtf.py:665-666).  The only exchange is the all-gather of the result tensor (north_star):

* ``GatherBuffer`` + ``ShardedFfdtf``: the B200 path.  Every rank holds the whole ``(total_windows, m, m, F)`` result;
  K5's finalize kernel writes each chunk of windows straight into the rank's own slot and ``gather_push_kernel``
  (csrc/gather_kernels.cu) streams the chunk to the peers over NVLink (plain stores to peer mappings, or one
  ``multimem.st`` through an NVSwitch multicast mapping) on a second stream while the next chunk computes.
* ``all_gather_windows``: the same exchange as one ``torch.distributed`` collective (NCCL on GPUs, gloo on CPU
  tensors in the tests) -- the baseline the push kernel is measured against."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_units(n_units: int, rank: int, world: int) -> range:
    """Contiguous block partition: ranks 0..(n_units % world)-1 get one extra unit."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def shard_by_cost(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time assignment for uneven units (tasks of different duration): returns, per rank,
    the unit indices it owns (deterministic)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    owned: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        owned[r].append(i)
        load[r] += costs[i]
    return [sorted(o) for o in owned]


def all_gather_windows(local, counts: Sequence[int]):
    """Gather per-rank result tensors ``(n_local, ...)`` (possibly different n_local) into ``(sum(counts), ...)``
    on every rank.  Equal counts use one ``all_gather_into_tensor``; ragged counts pad to the maximum."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    assert len(counts) == world
    tail = tuple(local.shape[1:])
    mx = max(counts)
    if all(c == mx for c in counts):
        out = torch.empty((world * mx,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    padded = torch.zeros((mx,) + tail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world * mx,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded)
    return torch.cat([buf[r * mx: r * mx + counts[r]] for r in range(world)], dim=0)


def unit_table(n_dyads: int, tasks: Sequence[str]) -> List[Tuple[int, str]]:
    """(dyad, task) units in the order run_pipeline visits them."""
    return [(d, t) for d in range(n_dyads) for t in tasks]


def bind_to_gpu_numa(device_index):
    """Pin the calling process to the CPUs local to GPU ``device_index`` (sysfs ``local_cpulist`` of its PCI function), so that
    pinned host buffers allocated afterwards (first touch) sit on the GPU's own NUMA node and the result stream of every rank
    crosses its own root complex instead of the socket interconnect -- what ``numactl --cpunodebind`` does per rank in a
    deployment.  Best effort: returns the CPU set, or None when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        addr = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{addr}/local_cpulist") as fh:
            text = fh.read().strip()
        cpus = set()
        for part in text.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


# --------------------------------------------------------------------------------------------- gather buffer
def window_layout(counts: Sequence[int]) -> List[int]:
    """Start offset (in windows) of every rank's slot in the gathered result: exclusive prefix sum of ``counts``."""
    out, acc = [], 0
    for c in counts:
        out.append(acc)
        acc += int(c)
    return out


def chunk_ranges(n_units: int, units_per_chunk: int, ramp: bool = False, ramp_down: bool = True) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) unit ranges of at most ``units_per_chunk`` units (the push granularity).  ``ramp``: the first chunks
    hold 1, 2, 4 units, so the exchange starts after a fraction of a full chunk's compute time, and (``ramp_down``) the last ones
    2 and 1, so little is left to push when the compute ends -- useful while the step is compute-bound (N <= 4); when the NVLink
    receive side bounds it (N = 8) the pushes queue up anyway and small last chunks only add copies."""
    if units_per_chunk < 1:
        raise ValueError("units_per_chunk must be >= 1")
    sizes: List[int] = []
    left = n_units
    if ramp:
        tail = [s for s in (2, 1) if s < units_per_chunk] if ramp_down else []
        k = 0
        while left > sum(tail) and (1 << k) < units_per_chunk:
            size = min(1 << k, left - sum(tail))
            sizes.append(size)
            left -= size
            k += 1
        while left > sum(tail):
            size = min(units_per_chunk, left - sum(tail))
            sizes.append(size)
            left -= size
        for s_ in tail:
            if left > 0:
                size = min(s_, left)
                sizes.append(size)
                left -= size
    while left > 0:
        size = min(units_per_chunk, left)
        sizes.append(size)
        left -= size
    out, lo = [], 0
    for size in sizes:
        out.append((lo, lo + size))
        lo += size
    return out


class GatherBuffer:
    """``numel`` float64 on every rank of the default process group, each rank able to store into every peer's copy.

    mode 'symm'  torch symmetric memory (``torch.distributed._symmetric_memory``): peer pointers + an NVSwitch
                 multicast pointer when the fabric offers one;
    mode 'ipc'   plain ``cudaMalloc`` buffers exchanged as CUDA IPC handles (hs_ipc_*), peer pointers only;
    mode 'auto'  'symm', falling back to 'ipc' when the symmetric allocation fails (same choice on every rank).
    world == 1 needs no peers: a plain tensor.
    """

    def __init__(self, numel: int, mode: str = "auto"):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        self.numel = int(numel)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.mode = None
        self.multicast_ptr = 0
        self.peer_ptrs: List[int] = []
        self._lib = _lib.load()
        self._ipc_opened: List[int] = []
        self._ipc_own = None
        self._symm = None
        dev = torch.device("cuda", torch.cuda.current_device())
        if self.world == 1:
            self.tensor = torch.empty(self.numel, dtype=torch.float64, device=dev)
            self.peer_ptrs = [self.tensor.data_ptr()]
            self.mode = "local"
            return
        if mode in ("auto", "symm"):
            ok = 1
            try:
                import torch.distributed._symmetric_memory as symm_mem
                t = symm_mem.empty(self.numel, dtype=torch.float64, device=dev)
                hdl = symm_mem.rendezvous(t, dist.group.WORLD)
                self.tensor, self._symm = t, hdl
                self.peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
                self.multicast_ptr = int(getattr(hdl, "multicast_ptr", 0) or 0)
            except Exception as exc:      # noqa: BLE001 -- any failure means "not available here"
                if mode == "symm":
                    raise
                self._symm_error = repr(exc)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                self.mode = "symm"
                return
            self._symm = None
            self.tensor = None
            self.multicast_ptr = 0
        # CUDA IPC
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(self._lib.hs_ipc_alloc(C.byref(ptr), self.numel * 8, handle), "hs_ipc_alloc")
        self._ipc_own = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw))
        self.peer_ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.peer_ptrs.append(self._ipc_own)
                continue
            q = C.c_void_p()
            _lib.check(self._lib.hs_ipc_open(C.c_char_p(h), C.byref(q)), "hs_ipc_open")
            self._ipc_opened.append(q.value)
            self.peer_ptrs.append(q.value)
        self.tensor = _tensor_from_ptr(self._ipc_own, self.numel, dev)
        self.mode = "ipc"

    @property
    def local_ptr(self) -> int:
        return self.peer_ptrs[self.rank]

    def remote_ptrs(self) -> List[int]:
        """Mapped base pointers of the OTHER ranks' buffers, in rotation order (rank+1, rank+2, ...): when every rank walks
        its list in step, each GPU receives from one sender at a time instead of all ranks converging on rank 0 first."""
        return [self.peer_ptrs[(self.rank + d) % self.world] for d in range(1, self.world)]

    def close(self):
        import torch
        torch.cuda.synchronize()
        for q in self._ipc_opened:
            self._lib.hs_ipc_close(q)
        self._ipc_opened = []
        self.tensor = None
        if self._ipc_own:
            self._lib.hs_ipc_free(self._ipc_own)
            self._ipc_own = None
        self._symm = None


def _tensor_from_ptr(ptr: int, numel: int, device):
    """float64 CUDA tensor viewing ``numel`` doubles at device pointer ``ptr`` (memory owned elsewhere)."""
    import torch

    class _Arr:
        __cuda_array_interface__ = {"shape": (numel,), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}

    return torch.as_tensor(_Arr(), device=device)


class ShardedFfdtf:
    """Windowed ffDTF of this rank's units with the result all-gathered chunk by chunk, overlapped with compute.

    ``x_all``: CUDA float64 ``(n_units_local, m, T)`` (all units the same length), windows ``starts`` (the same for every
    unit, ``_create_windows``), so unit u / window k is global window ``rank_offset + u * n_win + k`` of the gathered
    ``(total_windows, m, m, F)`` tensor.  Every rank must own the same number of units (pad with repeats otherwise).
    push: 'ce' (one peer copy per destination on the copy engines: measured fastest, 774 GB/s in per rank at N = 2, and it
    leaves every SM to the MVAR kernels), 'p2p' (gather_push_kernel: stores to every peer mapping, 650 GB/s with 16 CTAs),
    'multicast' (gather_push_kernel with multimem.st through the NVSwitch multicast mapping; needs mode 'symm'),
    'nccl' (ONE in-place all_gather_into_tensor after the last chunk: the unoverlapped baseline), 'none' (no exchange).
    """

    def __init__(self, n_units_local, m, T, window_size, starts, freqs, fs, p, units_per_chunk=5, push="ce", push_ctas=16,
                 buffer_mode="auto", ramp=True):
        import ctypes as C
        import numpy as np
        import torch
        import torch.distributed as dist
        from . import _lib
        self.torch, self.dist, self.C = torch, dist, C
        self.lib = _lib.load()
        self._check = _lib.check
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.m, self.T, self.W, self.p, self.fs = int(m), int(T), int(window_size), int(p), float(fs)
        self.n_units = int(n_units_local)
        st = np.asarray(starts, dtype=np.int64)
        self.n_win = int(st.size)
        self.F = int(np.asarray(freqs).size)
        self.per_win = self.m * self.m * self.F
        self.win_local = self.n_units * self.n_win
        self.total_windows = self.win_local * self.world
        self.rank_offset = self.rank * self.win_local
        self.chunks = chunk_ranges(self.n_units, units_per_chunk, ramp=ramp and self.world > 1, ramp_down=self.world <= 4)
        self.push = push
        self.push_ctas = int(push_ctas)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = GatherBuffer(self.total_windows * self.per_win, mode=buffer_mode)
        if push == "multicast" and not self.buf.multicast_ptr and self.world > 1:
            raise _lib.HsError("push='multicast' needs a symmetric-memory buffer with an NVSwitch multicast mapping")
        self.result = self.buf.tensor.view(self.total_windows, self.m, self.m, self.F)
        # element offsets of every local window: unit u, window k -> u * m * T + start_k  (channel stride T)
        offs = (np.arange(self.n_units, dtype=np.int64)[:, None] * (self.m * self.T) + st[None, :]).ravel()
        self.offsets = torch.from_numpy(offs).to(dev)
        self.freqs = torch.from_numpy(np.ascontiguousarray(freqs, dtype=np.float64)).to(dev)
        max_chunk_win = max(hi - lo for lo, hi in self.chunks) * self.n_win
        ws_bytes = int(self.lib.hs_mvar_ffdtf_ws_bytes(max_chunk_win, self.m, self.p, self.F))
        self.ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)      # chunks run back to back on one stream
        self.status = torch.zeros(self.win_local, dtype=torch.int32, device=dev)
        self.s_push = torch.cuda.Stream(device=dev)
        self.ev = [torch.cuda.Event() for _ in self.chunks]
        self._sync = torch.zeros(1, dtype=torch.int32, device=dev)
        remote = self.buf.remote_ptrs()
        self._n_remote = len(remote)
        self._remote = remote
        self.sm_total = torch.cuda.get_device_properties(dev).multi_processor_count
        # the push kernel's CTAs need SMs of their own next to K5 (one K5 CTA owns a whole SM's register file)
        self.sm_limit = self.sm_total - self.push_ctas if (self.world > 1 and push in ("p2p", "multicast")) else 0
        self._check(self.lib.hs_set_compute_sm_limit(self.sm_limit), "hs_set_compute_sm_limit")

    def _peer_array(self, elem_off):
        C = self.C
        arr = (C.c_void_p * max(self._n_remote, 1))()
        for i, base in enumerate(self._remote):
            arr[i] = base + 8 * elem_off
        return arr

    def step(self, x_all):
        """One pass over this rank's units; returns after ENQUEUEING everything (the caller synchronises).  On return of
        ``finish()`` every rank's slot of ``self.result`` holds that rank's windows."""
        torch = self.torch
        cur = torch.cuda.current_stream()
        assert tuple(x_all.shape) == (self.n_units, self.m, self.T) and x_all.is_contiguous()
        self.s_push.wait_stream(cur)
        for ci, (lo, hi) in enumerate(self.chunks):
            nw = (hi - lo) * self.n_win
            w0 = lo * self.n_win
            elem_off = (self.rank_offset + w0) * self.per_win
            self._check(self.lib.hs_mvar_ffdtf_f64(x_all.data_ptr(), self.offsets.data_ptr() + 8 * w0, self.T, nw, self.m, self.W, self.p,
                                                   self.freqs.data_ptr(), self.F, self.fs, self.buf.local_ptr + 8 * elem_off, None, None,
                                                   self.status.data_ptr() + 4 * w0, self.ws.data_ptr(), cur.cuda_stream),
                        "hs_mvar_ffdtf_f64")
            self.ev[ci].record(cur)
            if self.world == 1:
                continue
            self.s_push.wait_event(self.ev[ci])
            count = nw * self.per_win
            src = self.buf.local_ptr + 8 * elem_off
            if self.push == "multicast":
                self._check(self.lib.hs_gather_push_f64(src, count, self.buf.multicast_ptr + 8 * elem_off, None, 0, self.push_ctas,
                                                        self.s_push.cuda_stream), "hs_gather_push_f64")
            elif self.push == "p2p":
                self._check(self.lib.hs_gather_push_f64(src, count, None, self._peer_array(elem_off), self._n_remote, self.push_ctas,
                                                        self.s_push.cuda_stream), "hs_gather_push_f64")
            elif self.push == "ce":
                self._check(self.lib.hs_gather_push_ce(src, count, self._peer_array(elem_off), self._n_remote, self.s_push.cuda_stream),
                            "hs_gather_push_ce")
            elif self.push in ("nccl", "none"):
                pass      # 'nccl' (baseline): one in-place all_gather of the whole slot in finish(); 'none': compute only
            else:
                raise ValueError(f"unknown push mode {self.push!r}")

    def finish(self):
        """Join the push stream and cross the rank barrier: afterwards all slots of ``self.result`` are complete."""
        torch, dist = self.torch, self.dist
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.s_push)
        if self.world > 1:
            if self.push == "none":
                return
            if self.push == "nccl":
                flat = self.buf.tensor
                dist.all_gather_into_tensor(flat, flat[self.rank_offset * self.per_win:(self.rank_offset + self.win_local) * self.per_win])
            else:
                dist.all_reduce(self._sync)          # every rank's pushes were issued before its contribution: rank barrier

    def close(self):
        self.lib.hs_set_compute_sm_limit(0)
        self.result = None
        self.buf.close()
