"""Multi-GPU check of the symmetric pair shares (run under torchrun, one rank per GPU):
K iterations with N ranks (pair sums -> in-place NCCL reduce-scatter -> step -> all-gather) against
the single-rank plan on rank 0's GPU.  Prints the largest coordinate difference.
usage: python -m torch.distributed.run --nproc-per-node N tools/check_sym_ranks.py [n] [dim] [iters]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs, sharding

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = capi.Context(local, stream=stream.cuda_stream)
A = graphs.rgg(n, 10.0, seed=7)
n = A.shape[0]
x0 = capi.reference_uniform(23, n * dim).reshape(n, dim)
r0, r1, R, ld = sharding.row_block(n, world, rank)
plan = ctx.flat_plan(A, dim, capi.flat_params(), symmetric=(rank, world))
bufs = [torch.zeros(dim * ld, dtype=torch.float64, device=dev) for _ in range(2)]
sums = torch.zeros(dim * ld, dtype=torch.float64, device=dev)
plan.bind_coords(bufs[0].data_ptr(), bufs[1].data_ptr())
plan.bind_pair_sums(sums.data_ptr())
plan.upload(x0)
by_ptr = {b.data_ptr(): b for b in bufs}
for _ in range(iters):
    plan.launch_repulsion()
    sharding.reduce_scatter_pair_sums(dist, sums.view(dim, ld), rank, R)
    plan.launch_step()
    sharding.allgather_coords(dist, by_ptr[plan.next_ptr()].view(dim, ld), rank, R)
    plan.swap()
x = plan.download()
plan.close()
if rank == 0:
    os.environ["GE_NO_REORDER"] = "1"
    one = ctx.flat_plan(A, dim, capi.flat_params())
    one.upload(x0)
    one.iterate(iters)
    ref = one.download()
    one.close()
    print("ranks=%d n=%d dim=%d iterations=%d  max |x_ranks - x_single| = %.3e  (max |x| %.3f)"
          % (world, n, dim, iters, np.abs(x - ref).max(), np.abs(ref).max()), flush=True)
dist.barrier()
dist.destroy_process_group()
