"""Host-side logic of the N>1 path on CPU: world_size-2 (and 4) gloo processes build their rank views of one box and
exchange the corner keys exactly like Context.comm_init does; no GPU involved."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, kind="cart"):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from amps_b200 import mesh as meshmod, workload

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    decomp = "sfc" if kind == "sfc" else {2: (2, 1, 1), 4: (2, 2, 1)}[world]
    n_cells = (32, 32, 16)
    m = meshmod.uniform_periodic_box(n_cells, rank=rank, n_ranks=world, decomp=decomp)
    gathered = [None] * world
    dist.all_gather_object(gathered, m.corner_target_gkeys)
    lists = meshmod.shared_corner_lists(m, gathered)
    # symmetric exchange: what I send to r has the size r expects from me, in the same global-key order
    sizes = {r: len(u) for r, u in lists.items()}
    keys = {r: m.corner_gkey[u] for r, u in lists.items()}
    info = [None] * world
    dist.all_gather_object(info, (sizes, keys))
    ok = True
    for r, (sz, ky) in enumerate(info):
        if r == rank:
            continue
        ok &= sz.get(rank, 0) == sizes.get(r, 0)
        if rank in ky:
            ok &= bool((ky[rank] == keys[r]).all())
    # ownership is a partition of the real blocks; every particle of the global plasma has exactly one owner
    mg = meshmod.uniform_periodic_box(n_cells)
    x, v, w, sp, gcells = workload.maxwellian_box(mg, 2, seed=1)
    C = mg.cells_per_block
    pl = m.arrays["global_leaf_to_local"][gcells // C]
    mine = (pl >= 0) & (m.arrays["leaf_owner"][np.maximum(pl, 0)] == rank)
    counts = [None] * world
    dist.all_gather_object(counts, int(mine.sum()))
    ok &= sum(counts) == x.shape[1]
    # local geometry: boundary-layer blocks are adjacent to own blocks, ghost blocks carry their real image locally
    real = m.arrays["leaf_real"]
    fl = m.arrays["node_flags"][m.arrays["leaf_node"]]
    ghost = (fl & 2) != 0
    ok &= bool((real[ghost] >= 0).all()) and bool((real[~ghost] == -1).all())
    ok &= m.n_own_leaves == int(np.prod(np.array(n_cells) // 8)) // world
    # ---- the halo lists of the sharded field solve (Context.field_solver_init): every physical corner has exactly one primary rank,
    # what a rank sends to a peer is what the peer expects from it (same keys, same order), and one halo round of a vector that every
    # rank fills only on its primary corners reproduces the global vector on every copy a rank holds
    fg = [None] * world
    dist.all_gather_object(fg, (m.corner_gkey, m.corner_target_gkeys, m.center_gkey, m.center_own_gkeys))
    mask, hl = meshmod.field_halo_lists(m, fg)
    prim = [None] * world
    dist.all_gather_object(prim, m.corner_gkey[mask == 1])
    allp = np.concatenate(prim)
    ok &= len(allp) == len(np.unique(allp)) == mg.n_corners
    send = {r: (m.corner_gkey[l[0]], m.center_gkey[l[2]]) for r, l in hl.items()}
    recv = {r: (m.corner_gkey[l[1]], m.center_gkey[l[3]]) for r, l in hl.items()}
    ex = [None] * world
    dist.all_gather_object(ex, (send, recv))
    for r, (snd, rcv) in enumerate(ex):
        if r == rank:
            continue
        if rank in snd:  # what r sends to me == what I expect from r
            ok &= bool(np.array_equal(snd[rank][0], recv[r][0]) and np.array_equal(snd[rank][1], recv[r][1]))
        else:
            ok &= r not in recv or (len(recv[r][0]) == 0 and len(recv[r][1]) == 0)
    gvec = np.random.default_rng(5).standard_normal(mg.n_corners)
    pos = np.searchsorted(mg.corner_gkey, m.corner_gkey)
    loc = np.where(mask == 1, gvec[pos], np.nan)  # only the primary corners are filled
    out_msgs = {r: loc[l[0]] for r, l in hl.items()}
    msgs = [None] * world
    dist.all_gather_object(msgs, out_msgs)
    for r, l in hl.items():
        loc[l[1]] = msgs[r][rank]
    targets = np.searchsorted(m.corner_gkey, m.corner_target_gkeys)
    tiles = np.unique(m.arrays["leaf_corner_uid"].reshape(m.n_leaves, -1)[: m.n_own_leaves])
    tiles = tiles[tiles >= 0]
    ok &= bool(np.array_equal(loc[targets], gvec[pos][targets])) and not np.isnan(loc[tiles]).any()
    q.put((rank, bool(ok), sizes))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "cart"), (4, "cart"), (2, "sfc"), (4, "sfc")])
def test_rank_views_are_consistent(world, kind):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world + (10 if kind == "sfc" else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, kind)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    # 2 ranks split along x share two faces (x = 16 and the periodic seam): 2 * 32*16... corners
    if world == 2 and kind == "cart":
        assert all(sum(r[2].values()) == 2 * 32 * 16 for r in res), res


def test_space_filling_curve_chunks_follow_the_reference_rule():
    """decomp="sfc" (RedistributeParallelLoad, meshAMRgeneric.h:11905-11940): the chunks are contiguous on the Morton curve, in rank
    order, none empty; with the particle number as the load measure every rank's load is within one leaf of the mean; and the
    leaf that crosses a whole-number mark of the normalised cumulative load is the first of the next chunk."""
    sys.path.insert(0, ROOT)
    from amps_b200 import workload

    ppc = (64, 16, 8)
    world = 4
    views = [workload.amr_sphere_box((4, 4, 4), radii=(12.0, 6.0), rank=r, n_ranks=world, decomp="sfc",
                                     leaf_weight=lambda lev, lo, hi: ppc[lev]) for r in range(world)]
    g = workload.amr_sphere_box((4, 4, 4), radii=(12.0, 6.0))
    real = g.real_leaves()
    order = np.argsort(g.leaf_global[real])  # the global leaf numbering is the tree traversal = the Morton curve
    gid = g.leaf_global[real][order]
    wts = np.array([ppc[l] for l in g.leaf_level()[real][order]], dtype=np.float64)
    norm = wts.sum() / world
    owner = np.full(len(wts), -1)
    for r, m in enumerate(views):
        at = np.searchsorted(gid, m.leaf_global[: m.n_own_leaves])
        assert (gid[at] == m.leaf_global[: m.n_own_leaves]).all() and (owner[at] == -1).all()
        owner[at] = r
    assert (owner >= 0).all() and (np.diff(owner) >= 0).all() and len(np.unique(owner)) == world
    loads = np.array([wts[owner == r].sum() for r in range(world)])
    assert np.abs(loads - norm).max() <= wts.max(), loads
    cum = np.cumsum(wts / norm)
    for r in range(1, world):
        first = int(np.argmax(owner == r))
        assert cum[first] > r + 1e-8 and cum[first - 1] <= r + 1e-8
