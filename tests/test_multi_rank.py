"""world_size-2 run of the sharded evaluation on CPU (gloo): each rank evaluates its contiguous shard
of one batch (here with the CPU oracle standing in for the device, the host logic is what is under
test), reduces every instance to {f, max bound violation} and all-gathers the rows; the result must
equal the single-process evaluation of the whole batch."""
import os
import subprocess
import sys

import numpy as np

from etol_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import oracle_binding as ob
from etol_b200 import shard, workloads as W
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
B = 7                                   # deliberately not a multiple of the world size
wl = W.pm3d(batch=B, nnodes=12, ncyl=3)
per = shard.padded_shard(B, world)
lo, hi = shard.shard_range(B, world, rank)
assert hi - lo <= per
rows = np.zeros((per, 2))
if hi > lo:
    sub = wl.slice_batch(lo, hi)
    r = ob.Oracle(sub).eval(sub.x, want=("f", "g"), jac_mode=0, style=1, nthreads=1)
    viol = np.maximum(sub.gl - r["g"], r["g"] - sub.gu).max(axis=1).clip(min=0.0)
    rows[: hi - lo, 0], rows[: hi - lo, 1] = r["f"], viol
# ranks hold [per] rows each; shard_range gives the first `rem` ranks one more instance, so gather
# the padded blocks and pick the valid rows rank by rank
allrows = shard.gather_rows(torch.from_numpy(rows)).numpy().reshape(world, per, 2)
out = np.concatenate([allrows[r][: shard.shard_range(B, world, r)[1] - shard.shard_range(B, world, r)[0]]
                      for r in range(world)])
if rank == 0:
    np.save(sys.argv[2], out)
dist.destroy_process_group()
'''


def test_shard_ranges_cover_batch():
    for n in (0, 1, 7, 4096, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
            assert shard.padded_shard(n, world) >= max(h - l for l, h in spans)


def test_two_ranks_match_single_process(tmp_path):
    import oracle_binding as ob
    from etol_b200 import workloads as W
    script, out = tmp_path / "worker.py", tmp_path / "rows.npy"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", "29541", str(script), ROOT, str(out)],
                   check=True, env=env, timeout=600)
    got = np.load(out)
    wl = W.pm3d(batch=7, nnodes=12, ncyl=3)
    r = ob.Oracle(wl).eval(wl.x, want=("f", "g"), jac_mode=0, style=1, nthreads=1)
    viol = np.maximum(wl.gl - r["g"], r["g"] - wl.gu).max(axis=1).clip(min=0.0)
    assert got.shape == (7, 2)
    assert np.array_equal(got[:, 0], r["f"]) and np.array_equal(got[:, 1], viol)
