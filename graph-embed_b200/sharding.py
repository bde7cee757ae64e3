"""Row-block sharding of the flat ForceAtlas iteration across ranks (one process per GPU).

Rank r owns rows [r*R, min((r+1)*R, n)) with R = ld / world, where ld is the leading dimension of
the SoA coordinate buffers (n padded to the 256-entry column tile, hence divisible by 1, 2, 4, 8).
Every rank keeps the full coordinates; after each iteration the ranks exchange their freshly
written slices with one in-place all-gather per coordinate dimension (the reference's global
swing / traction sums are dead code, SURVEY.md section 0.3, so there is no all-reduce).

Symmetric plans (large graphs): the repulsion term is antisymmetric, so every unordered pair is
evaluated once, on one rank -- the upper triangle of (row block, column tile) units is cut into
equal shares (`pair_share`), each rank accumulates its pairs' contributions over the full length,
and one in-place reduce-scatter per dimension (`reduce_scatter_pair_sums`) hands every rank the
complete sums of its own rows before attraction + step."""


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


def pair_share(ld, world, rank, rows_per_block=1024, tile=256):
    """The share of the symmetric all-pairs sweep that rank `rank` evaluates (mirror of
    sym_layout in csrc/ge_flat_sym.cu).  The upper triangle is the list of (row block, column
    tile) units -- block g covers rows [g*rows_per_block, ...) and the tiles from its own first row
    to the end -- cut into `world` equal contiguous shares.  Returns
    [(row0, row1, tile_first, ntiles, tile_sym0)]: tiles below tile_sym0 lie inside the block's own
    rows (every ordered pair evaluated, row side only); tiles from tile_sym0 on are evaluated once
    and applied to both the rows and the columns."""
    ntile = ld // tile
    nblk = (ld + rows_per_block - 1) // rows_per_block
    per_block = [ntile - (g * rows_per_block) // tile for g in range(nblk)]
    total = sum(per_block)
    U0, U1 = total * rank // world, total * (rank + 1) // world
    out, prefix = [], 0
    for g in range(nblk):
        tf = (g * rows_per_block) // tile
        b0, b1 = prefix, prefix + per_block[g]
        prefix = b1
        lo, hi = max(b0, U0), min(b1, U1)
        if lo >= hi:
            continue
        row0, row1 = g * rows_per_block, min(ld, (g + 1) * rows_per_block)
        out.append((row0, row1, tf + (lo - b0), hi - lo, (row1 + tile - 1) // tile))
    return out


def reduce_scatter_pair_sums(dist, sums, rank, R):
    """sums: [dim, ld] tensor of this rank's raw pair sums over the full length.  Afterwards its
    columns [rank*R, (rank+1)*R) hold the sum over all ranks (in place; NCCL: recvbuff == sendbuff
    + rank * count).  gloo has no reduce-scatter: the CPU tests use an all-reduce instead."""
    for k in range(sums.shape[0]):
        if dist.get_backend() == "gloo":
            dist.all_reduce(sums[k])
        else:
            dist.reduce_scatter_tensor(sums[k, rank * R:(rank + 1) * R], sums[k])


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
