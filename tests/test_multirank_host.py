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


def test_aggregate_blocks_partition(graphs):
    from graph_embed_b200 import sharding
    As, Ps = graphs.coarsen(graphs.rmat(11, 8, seed=3), 0.25, min_coarse=30)
    for world in (1, 2, 4, 8):
        blocks = sharding.aggregate_blocks(As[0], Ps[0], world)
        assert blocks[0][0] == 0 and blocks[-1][1] == Ps[0].shape[0]
        assert all(b[1] == c[0] and b[0] <= b[1] for b, c in zip(blocks, blocks[1:]))
        s = np.diff(Ps[0].indptr).astype(float)
        cost = np.array([(s[b:e] ** 2).sum() for b, e in blocks])
        if world > 1:  # no rank carries more than its share plus the largest single aggregate
            assert cost.max() <= (s ** 2).sum() / world + (s ** 2).max() * 1.5 + Ps[0].shape[1]


def _sum_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import conftest  # noqa: F401
    from graph_embed_b200 import graphs, sharding
    O = conftest.ORACLE
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    As, Ps = graphs.coarsen(graphs.grid2d(14, 14), 0.25, min_coarse=10)
    A, P = As[0], Ps[0]
    m = P.shape[0]
    rng = np.random.default_rng(0)
    cA, rA = rng.normal(size=(m, 2)), rng.random(m) + 0.1
    full = O.multilevel_run(A, P, cA, rA, 2, O.multilevel_init(P, 2, 3), O.Params(iterations=10))
    b, e = sharding.aggregate_blocks(A, P, world)[rank]
    mine = np.zeros_like(full)            # what ge_multilevel_forceatlas_shard returns: own rows, zeros elsewhere
    rows = P.indices[P.indptr[b]:P.indptr[e]]
    mine[rows] = full[rows]
    t = torch.from_numpy(mine)
    dist.all_reduce(t)
    if rank == 0:
        np.save(out, np.stack([t.numpy(), full]))
    dist.barrier()
    dist.destroy_process_group()


def test_level_output_exchange_is_exact(tmp_path):
    """Sharded aggregates + one sum all-reduce of the level output == the unsharded level."""
    out = str(tmp_path / "lvl.npy")
    mp.spawn(_sum_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got, full = np.load(out)
    assert np.array_equal(got, full)


def _pair_sums_share(x, c, share, tile=256):
    """numpy restatement of one rank's share of the symmetric sweep: raw sums
    S_i += c_j (xi-xj)/dis^3 (rows), S_j -= c_i (xi-xj)/dis^3 (columns of symmetric tiles)."""
    ld, dim = x.shape
    S = np.zeros_like(x)
    for row0, row1, tf, nt, tsym in share:
        for t in range(tf, tf + nt):
            j0, j1 = t * tile, (t + 1) * tile
            d = x[row0:row1, None, :] - x[None, j0:j1, :]
            r2 = np.maximum((d * d).sum(-1), 1e-10)
            s = r2 ** -1.5
            S[row0:row1] += (d * (s * c[None, j0:j1])[..., None]).sum(1)
            if t >= tsym:
                S[j0:j1] -= (d * (s * c[row0:row1, None])[..., None]).sum(0)
    return S


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_pair_shares_cover_every_unordered_pair_once(world):
    """Summed over the ranks, the shares reproduce the ordered all-pairs sum of
    include/forceatlas.hpp:151-167 (before the c_i * repel factor)."""
    from graph_embed_b200 import sharding
    rng = np.random.default_rng(1)
    n, ld, dim = 2300, 2304, 2
    x = np.zeros((ld, dim))
    x[:n] = rng.uniform(-1, 1, (n, dim))
    c = np.zeros(ld)
    c[:n] = rng.integers(1, 9, n)
    d = x[:, None, :] - x[None, :, :]
    s = np.maximum((d * d).sum(-1), 1e-10) ** -1.5
    full = (d * (s * c[None, :])[..., None]).sum(1)
    shares = [sharding.pair_share(ld, world, r) for r in range(world)]
    units = sorted((row0, t) for sh in shares for row0, _, tf, nt, _ in sh for t in range(tf, tf + nt))
    assert len(units) == len(set(units))                      # no unit evaluated twice
    assert len(units) == sum(ld // 256 - r0 // 256 for r0 in range(0, ld, 1024))
    got = sum(_pair_sums_share(x, c, sh) for sh in shares)
    assert np.abs(got[:n] - full[:n]).max() < 1e-9 * np.abs(full[:n]).max()
    counts = [sum(nt for *_, nt, _ in sh) for sh in shares]
    assert max(counts) - min(counts) <= 1                     # equal shares


def _sym_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import conftest  # noqa: F401
    from graph_embed_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(2)
    n, ld, dim = 2000, 2048, 3
    x = np.zeros((ld, dim))
    x[:n] = rng.uniform(-1, 1, (n, dim))
    c = np.zeros(ld)
    c[:n] = rng.integers(1, 9, n)
    S = _pair_sums_share(x, c, sharding.pair_share(ld, world, rank))
    sums = torch.from_numpy(np.ascontiguousarray(S.T))       # [dim, ld] like the device buffer
    R = ld // world
    sharding.reduce_scatter_pair_sums(dist, sums, rank, R)
    own = sums[:, rank * R:(rank + 1) * R].clone()
    gathered = [torch.zeros_like(own) for _ in range(world)]
    dist.all_gather(gathered, own)
    if rank == 0:
        np.save(out, torch.cat(gathered, dim=1).numpy().T)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_symmetric_pair_sums(tmp_path):
    """Two ranks, each with half of the unordered pairs: after the exchange every rank's own
    rows hold the complete sums."""
    out = str(tmp_path / "sums.npy")
    mp.spawn(_sym_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    rng = np.random.default_rng(2)
    n, ld, dim = 2000, 2048, 3
    x = np.zeros((ld, dim))
    x[:n] = rng.uniform(-1, 1, (n, dim))
    c = np.zeros(ld)
    c[:n] = rng.integers(1, 9, n)
    d = x[:, None, :] - x[None, :, :]
    s = np.maximum((d * d).sum(-1), 1e-10) ** -1.5
    full = (d * (s * c[None, :])[..., None]).sum(1)
    got = np.load(out)
    assert np.abs(got[:n] - full[:n]).max() < 1e-9 * np.abs(full[:n]).max()


def _attract_gravity_rows(A, x, c, r0, r1, attract=1.0, gravity=1.0):
    """include/forceatlas.hpp:169-211 for rows [r0, r1) with the default options: attraction
    (xj - xi) * attract * a_ij and gravity -x/|x| * gravity * (deg + 1)."""
    F = np.zeros((r1 - r0, x.shape[1]))
    for i in range(r0, r1):
        for e in range(A.indptr[i], A.indptr[i + 1]):
            j = A.indices[e]
            F[i - r0] += (x[j] - x[i]) * (attract * A.data[e])
        F[i - r0] -= x[i] / np.sqrt((x[i] * x[i]).sum()) * (gravity * c[i])
    return F


def _sym_iteration_worker(rank, world, port, iters, out):
    """The symmetric multi-rank iteration end to end with numpy standing in for the kernels:
    pair shares -> exchange of the pair sums -> attraction/gravity/step on the own row block ->
    in-place all-gather of the coordinates."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import conftest  # noqa: F401
    from graph_embed_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, z = load_flat_golden()
    n, dim = A.shape[0], 2
    r0, r1, R, ld = sharding.row_block(n, world, rank)
    c = np.zeros(ld)
    c[:n] = np.asarray(A.sum(axis=1)).ravel() + 1.0
    cur = torch.zeros(dim, ld, dtype=torch.float64)
    cur[:, :n] = torch.from_numpy(z["x0_d2"].T.copy())
    nxt = cur.clone()
    fprev = np.zeros((r1 - r0, dim))
    share = sharding.pair_share(ld, world, rank, rows_per_block=64, tile=16)  # several units at n = 144
    for _ in range(iters):
        x = np.ascontiguousarray(cur.numpy().T)                      # [ld, dim]
        S = _pair_sums_share(x, c, share, tile=16)
        sums = torch.from_numpy(np.ascontiguousarray(S.T))
        sharding.reduce_scatter_pair_sums(dist, sums, rank, R)
        F = sums[:, r0:r1].numpy().T * c[r0:r1, None] * 1.0          # * c_i * repel
        F = F + _attract_gravity_rows(A, x[:n], c, r0, r1)
        nxt[:, r0:r1] = torch.from_numpy(_step_rows(x[r0:r1], F, fprev).T.copy())
        fprev = F.copy()
        sharding.allgather_coords(dist, nxt, rank, R)
        cur, nxt = nxt, cur
    if rank == 0:
        np.save(out, cur[:, :n].numpy().T)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_symmetric_iterations_follow_the_reference(tmp_path, oracle):
    """Two iterations of the symmetric two-rank scheme land on the compiled reference's golden
    positions up to summation order (unordered pairs, partial sums per rank)."""
    iters = 2
    out = str(tmp_path / "sym_coords.npy")
    mp.spawn(_sym_iteration_worker, args=(2, _free_port(), iters, out), nprocs=2, join=True)
    _, z = load_flat_golden()
    ref = z["x_d2_k%d" % iters]
    assert np.abs(np.load(out) - ref).max() < 1e-9 * np.abs(ref).max()
