"""Host-side multi-rank logic on CPU: gloo, world_size 2 (no GPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from hyperscanning_signal_analysis_b200 import sharding
from hyperscanning_signal_analysis_b200.eeg_alpha_ibi_ffdtf import window_starts, create_windows


def test_shard_units_partition():
    for n in (0, 1, 7, 64, 192, 599):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                seen += list(sharding.shard_units(n, r, world))
            assert seen == list(range(n))
            sizes = [len(sharding.shard_units(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_units(4, 2, 2)
    owned = sharding.shard_by_cost([600, 120, 120, 300, 300, 60], 2)
    assert sorted(owned[0] + owned[1]) == list(range(6))
    loads = [sum([600, 120, 120, 300, 300, 60][i] for i in o) for o in owned]
    assert abs(loads[0] - loads[1]) <= 120
    assert sharding.unit_table(2, ["SECORE", "MOVIE", "TALK"])[4] == (1, "MOVIE")


def test_window_contract_matches_reference_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "window_starts.npz"))
    for key in g.files:
        if key == "versions":
            continue
        T, nw, ws = key.split("_")
        T, nw = int(T[1:]), int(nw[1:])
        ws = None if ws[1:] == "None" else int(ws[1:])
        starts, size = window_starts(T, nw, ws)
        assert starts.tolist() == g[key][:-1].tolist() and size == int(g[key][-1])
    sig = np.arange(40.0).reshape(2, 20)
    wl = create_windows(sig, 3, 10)
    assert [w[0, 0] for w in wl] == [0, 5, 10] and all(w.shape == (2, 10) for w in wl)
    for bad in ((1000, 3, None), (1000, 3, 100), (100, 3, 200), (100, 60, 99)):
        with pytest.raises(ValueError):
            window_starts(*bad)


WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["HS_ROOT"])
import torch, torch.distributed as dist
from hyperscanning_signal_analysis_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
units = sharding.unit_table(3, ["SECORE", "MOVIE", "TALK"])           # 9 units over 2 ranks: 5 + 4
mine = sharding.shard_units(len(units), rank, world)
counts = [len(sharding.shard_units(len(units), r, world)) for r in range(world)]
# stand-in for the per-unit result: (n_local, 2, 2, 3) filled with the unit index
local = torch.stack([torch.full((2, 2, 3), float(u), dtype=torch.float64) for u in mine]) if len(mine) else torch.empty((0, 2, 2, 3), dtype=torch.float64)
out = sharding.all_gather_windows(local, counts)
assert out.shape == (9, 2, 2, 3), out.shape
assert torch.equal(out[:, 0, 0, 0], torch.arange(9, dtype=torch.float64)), out[:, 0, 0, 0]
# equal-count path
out2 = sharding.all_gather_windows(torch.full((4, 3), float(rank), dtype=torch.float64), [4, 4])
assert out2.shape == (8, 3) and out2[:4].eq(0).all() and out2[4:].eq(1).all()
# slot layout of the gathered result and the push granularity
assert sharding.window_layout([5, 4, 0, 3]) == [0, 5, 9, 9]
assert sharding.chunk_ranges(24, 5) == [(0, 5), (5, 10), (10, 15), (15, 20), (20, 24)]
assert sharding.chunk_ranges(0, 5) == []
assert sharding.chunk_ranges(24, 5, ramp=True) == [(0, 1), (1, 3), (3, 7), (7, 12), (12, 17), (17, 21), (21, 23), (23, 24)]
assert sharding.chunk_ranges(24, 5, ramp=True, ramp_down=False) == [(0, 1), (1, 3), (3, 7), (7, 12), (12, 17), (17, 22), (22, 24)]
for n_u in range(0, 40):
    for upc in (1, 2, 5):
        cr = sharding.chunk_ranges(n_u, upc, ramp=True)
        assert [a for a, _ in cr] == [0] * (n_u > 0) + [b for _, b in cr][:-1] and (not cr or cr[-1][1] == n_u)
        assert all(0 < b - a <= upc for a, b in cr)
# cfg3: 64 dyads x 3 tasks split evenly over 1/2/4/8 ranks, whole dyads per rank
u3 = sharding.unit_table(64, ["SECORE", "MOVIE", "TALK"])
for wsz in (1, 2, 4, 8):
    got = [sharding.shard_units(len(u3), r, wsz) for r in range(wsz)]
    assert all(len(g) == len(u3) // wsz for g in got) and sum(len(g) for g in got) == 192
    for g in got:
        assert {u3[i][0] for i in g} & {u3[i][0] for h in got if h is not g for i in h} == set()
# max-over-ranks timing reduction used by bench.py
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == float(world)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, HS_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2
