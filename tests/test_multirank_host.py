"""World-size-2 test of the multi-GPU host logic on CPU (gloo): row-block sharding + the per-
iteration in-place coordinate all-gather reproduce the unsharded oracle bit for bit.  The per-row
arithmetic is the oracle's (the kernels are covered by the -m gpu tests); what is exercised here
is the partition, the SoA slice layout and the collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_flat_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _step_rows(x, f, fprev, ks=0.1, ksmax=1.0, gs=1.0):
    """include/forceatlas.hpp:214-261 for a block of rows, same operation order as the oracle."""
    d = fprev - f
    sw = np.zeros(len(x))
    tf = np.zeros(len(x))
    for k in range(x.shape[1]):
        sw = sw + d[:, k] * d[:, k]
        tf = tf + f[:, k] * f[:, k]
    swing, total = np.sqrt(sw), np.sqrt(tf)
    speed = ks * gs / (1 + gs * np.sqrt(swing))
    with np.errstate(divide="ignore"):
        cons = ksmax / total
    speed = np.where(speed > cons, cons, speed)
    return f * speed[:, None] + x


def _worker(rank, world, port, iters, dim, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import conftest  # noqa: F401  (loads the package and the oracle)
    from graph_embed_b200 import sharding
    O = conftest.ORACLE
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, z = load_flat_golden()
    n = A.shape[0]
    r0, r1, R, ld = sharding.row_block(n, world, rank)
    cur = torch.zeros(dim, ld, dtype=torch.float64)
    cur[:, :n] = torch.from_numpy(z["x0_d%d" % dim].T.copy())
    nxt = cur.clone()
    fprev = np.zeros((r1 - r0, dim))
    for _ in range(iters):
        x = cur[:, :n].numpy().T.copy()
        F, _ = O.flat_forces(A, dim, x, rows=(r0, r1))
        nxt[:, r0:r1] = torch.from_numpy(_step_rows(x[r0:r1], F[r0:r1], fprev).T.copy())
        fprev = F[r0:r1].copy()
        sharding.allgather_coords(dist, nxt, rank, R)
        cur, nxt = nxt, cur
    if rank == 0:
        np.save(out, cur[:, :n].numpy().T)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dim", [2, 3])
def test_two_rank_row_blocks_match_unsharded_oracle(tmp_path, oracle, dim):
    iters = 5
    out = str(tmp_path / "coords.npy")
    mp.spawn(_worker, args=(2, _free_port(), iters, dim, out), nprocs=2, join=True)
    _, z = load_flat_golden()
    assert np.array_equal(np.load(out), z["x_d%d_k%d" % (dim, iters)])


def test_row_block_partition_covers_every_row():
    from graph_embed_b200 import sharding
    for n in (1, 255, 256, 257, 499_920, 500_000):
        for world in (1, 2, 4, 8):
            blocks = [sharding.row_block(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(b[1] == c[0] for b, c in zip(blocks, blocks[1:]))
            assert all(b[2] * world == b[3] and b[3] % 256 == 0 for b in blocks)
