"""SURVEY 8f row f3, restart half, against the reference's OWN restart code (oracle/_ref/libref_pic.so):

  PIC::Restart::SaveParticleData (reference)  ->  amps_gpu_restart_read   : the device store holds exactly the reference's particles
  amps_gpu_restart_save                       ->  PIC::Restart::ReadParticleData (reference, after deleting its particles):
                                                  the reference holds exactly the particles it had before

"exactly" = per cell the same multiset of (species, x, v, weight correction) bit for bit.  Runs where the library was built
(it travels to the GPU box as a built .so)."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from amps_b200 import _capi, api
from oracle.ref_pic import ref_pic


def canon(cells, x, v, w, sp):
    o = np.lexsort((w.view(np.int64), v[2].view(np.int64), v[1].view(np.int64), v[0].view(np.int64), x[2].view(np.int64), x[1].view(np.int64),
                    x[0].view(np.int64), sp, cells))
    return cells[o], x[:, o], v[:, o], w[o], sp[o]


@pytest.mark.gpu
@pytest.mark.skipif(not ref_pic.available(), reason="oracle/_ref/libref_pic.so not built")
def test_restart_files_round_trip_through_the_reference():
    from tests import ref_ecsim_case as rc

    c = rc.case()
    r, m, cfg = c["refpic"], c["mesh"], c["cfg"]
    lib = r.lib
    L = np.zeros(7, dtype=np.int64)
    lib.ref_pic_record_layout(ref_pic._p(L))
    lay = _capi.AosLayout()
    lay.stride, lay.off_species, lay.off_v, lay.off_x, lay.off_w, lay.off_next, lay.off_prev = (int(v) for v in L)
    lay.off_mu, lay.off_vpar = -1, -1
    idb = lib.ref_pic_node_id_bytes()
    ids_ref = np.zeros((r.n_blocks, idb), dtype=np.uint8)
    lib.ref_pic_node_ids(ref_pic._p(ids_ref))
    # node id of every leaf of this repo's mesh: the reference block at the same place (ghost leaves: never written, any id)
    lx = m.leaf_xmin()
    ids = np.full((m.n_leaves, idb), 0xFF, dtype=np.uint8)
    b2l = {}
    for b in range(r.n_blocks):
        d = np.abs(lx - r.bxmin[b]).max(axis=1)
        k = int(np.argmin(d))
        if d[k] == 0.0 and r.ghost[b] == 0:
            ids[k] = ids_ref[b]
            b2l[b] = k
    header_bytes = 43 + 8 + 8 * r.n_species
    tmp = tempfile.mkdtemp(prefix="restart_")
    fa, fb = os.path.join(tmp, "a.restart"), os.path.join(tmp, "b.restart")
    p0 = r.particles()
    Cb = m.cells_per_block
    cells0 = np.array([b2l[b] for b in p0["block"]], dtype=np.int64) * Cb + p0["cell"]
    ref0 = canon(cells0, p0["x"], p0["v"], p0["w"], p0["species"])
    with ref_pic.quiet():
        lib.ref_pic_save_restart(fa.encode())
    g = api.Context(cfg, m)
    n = g.restart_read(fa, header_bytes, ids, lay)
    assert n == p0["x"].shape[1]
    d = g.particles_download()
    got = canon(d["cells"].astype(np.int64), d["x"], d["v"], d["w"], d["species"].astype(np.int32))
    for a, b in zip(got, ref0):
        assert np.array_equal(a, b)
    # and back: the header is the reference's own (marker, ParticleDataLength, weights)
    with open(fa, "rb") as f:
        header = f.read(header_bytes)
    assert g.restart_save(fb, header, ids, lay) == n
    g.close()
    lib.ref_pic_read_restart.restype = C.c_long
    with ref_pic.quiet():
        n_back = lib.ref_pic_read_restart(fb.encode())
    assert n_back == n
    p1 = r.particles()
    cells1 = np.array([b2l[b] for b in p1["block"]], dtype=np.int64) * Cb + p1["cell"]
    back = canon(cells1, p1["x"], p1["v"], p1["w"], p1["species"])
    for a, b in zip(back, ref0):
        assert np.array_equal(a, b)
