"""embed() with the multilevel levels sharded over ranks (sharding.embed_sharded: aggregates are
independent, one sum all-reduce of each level's output; the coarsest solve is replicated).
Run under torchrun, one rank per GPU; rank 0 prints one JSON line and checks the result against
the unsharded ge_embed call (bit-identical by construction).
usage: python -m torch.distributed.run --nproc-per-node N tools/run_embed_sharded.py [n] [dim]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs, sharding

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
ctx = capi.Context(local)
A = graphs.rgg(n, 10.0, seed=3)
As, Ps = graphs.coarsen(A, 0.25, min_coarse=64)


class One:
    @staticmethod
    def all_reduce(t):
        return t


comm = dist if world > 1 else One
sharding.embed_sharded(ctx, comm, As[-3:], Ps[-2:], dim, seed=5, rank=rank, world=world, device=dev,
                       coarse_iterations=100)  # warm-up
walls = []
for rep in range(2):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.time()
    x = sharding.embed_sharded(ctx, comm, As, Ps, dim, seed=5, rank=rank, world=world, device=dev)
    torch.cuda.synchronize()
    walls.append(time.time() - t)
if rank == 0:
    t = time.time()
    ref, st = ctx.embed(As, Ps, dim, seed=5)
    t_one = time.time() - t
    print(json.dumps({"n": A.shape[0], "nnz": int(A.nnz), "dim": dim, "levels": [a.shape[0] for a in As],
                      "ranks": world, "embed_sharded_wall_s": min(walls), "ge_embed_single_call_wall_s": t_one,
                      "coarse_ms_single": st["coarse_ms"], "levels_ms_single": st["levels_ms"],
                      "identical_to_single_call": bool(np.array_equal(x, ref))}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
