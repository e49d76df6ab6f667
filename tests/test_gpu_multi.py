"""Two ranks on two GPUs (NCCL): the sharded evaluation with the fused summary + all-gather over NVLink
peer memory (ecuda_summarize_allgather) must give exactly what the single-GPU summary + NCCL
all_gather gives. Skipped on boxes with fewer than two GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from etol_b200 import capi, shard, workloads as W
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 33
NN = int(sys.argv[3])
wl = W.pm3d(batch=B, nnodes=NN, ncyl=3, seed=W.SEED + rank)   # every rank its own shard
ev = capi.Evaluator(wl, device=local)
x = torch.from_numpy(wl.x).to(dev)
f = torch.empty(B, dtype=torch.float64, device=dev)
g = torch.empty((B, ev.ncons), dtype=torch.float64, device=dev)
# one explicit stream for everything: a NULL stream pointer would select the handle's own non-blocking
# stream (include/ecuda.h), which is not ordered with torch's current stream
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
st = stream.cuda_stream
assert st != 0
ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), None, capi.JAC_EXACT, capi.MEM_DEVICE, st)
summ = torch.empty((B, 2), dtype=torch.float64, device=dev)
ev.summarize_ptr(f.data_ptr(), g.data_ptr(), summ.data_ptr(), st)
ref = shard.gather_rows(summ)                                   # NCCL all_gather
buf = symm.empty((world * B, 2), dtype=torch.float64, device=dev)
buf.fill_(float("nan"))
hdl = symm.rendezvous(buf, dist.group.WORLD)
hdl.barrier(channel=0)                                          # everyone has cleared its buffer
ev.summarize_allgather_ptr(f.data_ptr(), g.data_ptr(), [int(p) for p in hdl.buffer_ptrs], rank, st)
hdl.barrier(channel=0)
torch.cuda.synchronize()
ok = torch.equal(buf, ref)
# evaluation and exchange fused in one kernel (ecuda_eval_allgather), both Jacobian modes
jac = torch.empty((B, ev.nnz), dtype=torch.float64, device=dev)
for mode in (capi.JAC_FD, capi.JAC_EXACT):
    buf.fill_(float("nan"))
    torch.cuda.synchronize()
    hdl.barrier(channel=0)
    ev.eval_allgather_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode,
                          [int(p) for p in hdl.buffer_ptrs], rank, st)
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    ok = ok and torch.equal(buf, ref)
# the library's own flag barrier in place of the symmetric-memory handle's
flags = symm.empty(64, dtype=torch.int64, device=dev)
flags.zero_()
fh = symm.rendezvous(flags, dist.group.WORLD)
torch.cuda.synchronize()
fh.barrier(channel=0)
fptrs = [int(p) for p in fh.buffer_ptrs]
for step in (1, 2, 3):
    buf.fill_(float("nan"))
    ev.peer_barrier_ptr(fptrs, rank, 2 * step - 1, st)      # everyone has cleared its buffer
    ev.eval_allgather_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), capi.JAC_FD,
                          [int(p) for p in hdl.buffer_ptrs], rank, st)
    ev.peer_barrier_ptr(fptrs, rank, 2 * step, st)
    torch.cuda.synchronize()
    ok = ok and torch.equal(buf, ref)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    open(sys.argv[2], "w").write("ok" if int(flag.item()) == 1 else "mismatch")
ev.close()
dist.destroy_process_group()
'''


@pytest.mark.gpu
@pytest.mark.parametrize("nnodes", [17, 40])  # 17: round-1 kernels (k_eval_rows); 40: the N-specialised k_rows_n
def test_fused_allgather_matches_nccl(tmp_path, nnodes):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip(f"multi-GPU parity of the fused exchange needs 2 GPUs; this box exposes {ngpu} "
                    "(the same check runs inside `bench.py --gpus N`, field gather_parity, and on one GPU in "
                    "test_gpu_parity.py::test_fused_summary_equals_separate_summary)")
    script, out = tmp_path / "worker.py", tmp_path / "result.txt"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", str(29547 + nnodes), str(script), ROOT, str(out),
                    str(nnodes)],
                   check=True, env=env, timeout=600)
    assert out.read_text() == "ok"
