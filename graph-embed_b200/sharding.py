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
