"""Row-block sharding of the flat ForceAtlas iteration across ranks (one process per GPU).

Rank r owns rows [r*R, min((r+1)*R, n)) with R = ld / world, where ld is the leading dimension of
the SoA coordinate buffers (n padded to the 256-entry column tile, hence divisible by 1, 2, 4, 8).
Every rank keeps the full coordinates; after each iteration the ranks exchange their freshly
written slices with one in-place all-gather per coordinate dimension (the only collective of the
path: the reference's global swing / traction sums are dead code)."""


def padded_ld(n, tile=256):
    return ((max(n, 1) + tile - 1) // tile) * tile


def row_block(n, world, rank, tile=256):
    """-> (row_begin, row_end, R, ld)"""
    ld = padded_ld(n, tile)
    if ld % world:
        raise ValueError("world size %d does not divide the padded row count %d" % (world, ld))
    R = ld // world
    return min(n, rank * R), min(n, (rank + 1) * R), R, ld


def allgather_coords(dist, nxt, rank, R):
    """nxt: [dim, ld] tensor holding this rank's new positions in columns [rank*R, (rank+1)*R).
    In-place all-gather per dimension (NCCL: sendbuff == recvbuff + rank * count)."""
    for k in range(nxt.shape[0]):
        dist.all_gather_into_tensor(nxt[k], nxt[k, rank * R:(rank + 1) * R])


def aggregate_blocks(A, P_T, world):
    """Contiguous aggregate ranges of (nearly) equal cost, one per rank, for the per-aggregate
    solver: cost(a) = s_a^2 ordered pairs + the CSR entries of its members' rows (SURVEY.md
    section 8e).  -> [(begin, end)] * world"""
    import numpy as np
    s = np.diff(P_T.indptr).astype(np.float64)
    row_nnz = np.diff(A.indptr).astype(np.float64)
    member_nnz = np.add.reduceat(row_nnz[P_T.indices], P_T.indptr[:-1]) if P_T.shape[0] else s
    member_nnz = np.where(s > 0, member_nnz, 0.0)
    cost = np.concatenate([[0.0], np.cumsum(s * s + member_nnz)])
    cuts = [int(np.searchsorted(cost, cost[-1] * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, P_T.shape[0]
    cuts = np.maximum.accumulate(cuts)
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def embed_sharded(ctx, dist, As, P_Ts, dim, seed, rank, world, device=None, precision=0,
                  coarse_iterations=100000, level_iterations=100):
    """partition::embed (src/embed.cpp:561-796) with the multilevel levels sharded over ranks:
    the coarsest flat solve is replicated (n ~ 30-100: replicas only), every rank computes the
    radii on its host, solves its range of aggregates (ge_multilevel_forceatlas_shard) and the
    level's coordinates are exchanged with one sum all-reduce (foreign rows are exact zeros).
    `seed` must be the same non-zero value on all ranks."""
    import numpy as np
    import torch
    from . import capi
    assert seed != 0
    L = len(P_Ts)
    n = As[L].shape[0]
    x = capi.reference_uniform(seed, n * dim).reshape(n, dim)
    coords = ctx.flat_forceatlas(As[L], dim, x, capi.flat_params(iterations=coarse_iterations,
                                                                 precision=precision))
    r_Ac = coords_Ac = None
    for l in range(L - 1, -1, -1):
        if r_Ac is None:
            coords_A, r_A = capi.level_radii(coords, dim)
        else:
            coords_A, r_A = capi.level_radii(coords, dim, As[l + 1], P_Ts[l + 1], coords_Ac, r_Ac)
        blocks = aggregate_blocks(As[l], P_Ts[l], world)
        part = ctx.multilevel_forceatlas(
            As[l], P_Ts[l], coords_A, r_A, dim,
            capi.multilevel_params(iterations=level_iterations, precision=precision, seed=seed),
            aggregates=blocks[rank])
        if world > 1:
            t = torch.from_numpy(part)
            if device is not None:
                t = t.to(device)
            dist.all_reduce(t)
            part = t.cpu().numpy()
        coords, r_Ac, coords_Ac = part, r_A, coords_A
    return coords
