"""NVLink result-exchange probe (run under torchrun, one rank per GPU): pushes a 2 GiB chunk from every rank to every peer
with each mechanism of sharding.GatherBuffer / hs_gather_push_* and prints achieved GB/s per rank (bytes RECEIVED per rank /
time, max over ranks).  Usage: python -m torch.distributed.run --nproc-per-node N tools/gather_probe.py [buffer_mode]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from hyperscanning_signal_analysis_b200 import _lib, sharding

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
lib = _lib.load()
mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
per_rank = 1 << 28                                   # doubles per rank slot: 2 GiB
buf = sharding.GatherBuffer(per_rank * world, mode=mode)
res = {"world": world, "buffer": buf.mode, "multicast": bool(buf.multicast_ptr), "chunk_bytes": per_rank * 8, "gbs_in_per_rank": {}}
if getattr(buf, "_symm_error", None):
    res["symm_error"] = buf._symm_error
off = rank * per_rank
buf.tensor[off:off + per_rank].normal_()
src = buf.local_ptr + 8 * off
remote = buf.remote_ptrs()
arr = (C.c_void_p * len(remote))(*[p + 8 * off for p in remote])
st = torch.cuda.current_stream().cuda_stream
sync = torch.zeros(1, dtype=torch.int32, device="cuda")


def run(name, fn, reps=3):
    fn()
    dist.all_reduce(sync)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    dist.all_reduce(sync)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res["gbs_in_per_rank"][name] = (world - 1) * per_rank * 8 * reps / (float(ms.item()) * 1e-3) * 1e-9


for ctas in (2, 4, 8, 16, 32):
    run(f"p2p_stores_{ctas}ctas", lambda: _lib.check(lib.hs_gather_push_f64(src, per_rank, None, arr, len(remote), ctas, st), "p2p"))
if buf.multicast_ptr:
    for ctas in (1, 2, 4, 8, 16):
        run(f"multicast_{ctas}ctas", lambda: _lib.check(lib.hs_gather_push_f64(src, per_rank, buf.multicast_ptr + 8 * off, None, 0, ctas, st), "mc"))
run("copy_engines", lambda: _lib.check(lib.hs_gather_push_ce(src, per_rank, arr, len(remote), st), "ce"))
run("nccl_all_gather_in_place", lambda: dist.all_gather_into_tensor(buf.tensor, buf.tensor[off:off + per_rank]))
# correctness of the last mechanism-independent state: every slot r must equal what rank r generated (checksum exchange)
sums = torch.stack([buf.tensor[r * per_rank:(r + 1) * per_rank].sum() for r in range(world)])
lo, hi = sums.clone(), sums.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
res["slots_identical_on_all_ranks"] = bool(torch.equal(lo, hi))
if rank == 0:
    print(json.dumps(res))
buf.close()
dist.destroy_process_group()
