"""N>1 host logic on CPU: two gloo ranks build their partitions with meshkit, exchange face-neighbour
element data exactly as the NCCL path does (one contiguous element-major message per peer, packed from
send_elems) and check every received halo element against the global field.  Also checks the global
face bookkeeping of the partition against the serial mesh."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tps_b200

N3 = (6, 5, 4)
ND, NF = 8, 3  # dofs per element and fields used for the synthetic payload


def _field(gid):
    """Deterministic per-element payload [len(gid), NF, ND]."""
    g = gid.astype(np.float64)[:, None, None]
    f = np.arange(NF, dtype=np.float64)[None, :, None]
    n = np.arange(ND, dtype=np.float64)[None, None, :]
    return 1000.0 * g + 10.0 * f + n + 0.5


def _worker(rank, world, port, procs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = tps_b200.cartesian_hex_partition(N3, procs, rank)
        ne = part["num_elems"]
        local = _field(part["elem_gid"][:ne])
        reqs, recv_bufs = [], []
        for p, peer in enumerate(part["nbr_rank"]):
            s0, s1 = part["send_offset"][p], part["send_offset"][p + 1]
            r0, r1 = part["recv_offset"][p], part["recv_offset"][p + 1]
            sb = torch.from_numpy(np.ascontiguousarray(local[part["send_elems"][s0:s1]]))  # pack
            rb = torch.empty((r1 - r0, NF, ND), dtype=torch.float64)
            recv_bufs.append((r0, r1, rb))
            reqs.append(dist.isend(sb, int(peer)))
            reqs.append(dist.irecv(rb, int(peer)))
        for r in reqs:
            r.wait()
        halo = np.zeros((part["num_nbr_elems"], NF, ND))
        for r0, r1, rb in recv_bufs:
            halo[r0:r1] = rb.numpy()
        ok = np.array_equal(halo, _field(part["elem_gid"][ne:]))
        el2 = part["face_el2"]
        stats = dict(rank=rank, ok=bool(ok), ne=ne, nh=part["num_nbr_elems"],
                     local_faces=int(((el2 >= 0) & (el2 < ne)).sum()), shared=int((el2 >= ne).sum()),
                     bdr=int((el2 < 0).sum()), el1_local=bool((part["face_el1"] < ne).all()))
        q.put(stats)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("procs", [(2, 1, 1), (1, 1, 2)])
def test_two_rank_halo_exchange_gloo(lib_built, procs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, procs, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    serial = tps_b200.cartesian_hex_mesh(*N3)
    assert all(r["ok"] and r["el1_local"] and r["bdr"] == 0 for r in res)
    assert sum(r["ne"] for r in res) == N3[0] * N3[1] * N3[2]
    # every global face is either local to one rank or shared by exactly two
    assert sum(r["local_faces"] for r in res) + sum(r["shared"] for r in res) // 2 == len(serial["face_el1"])


def test_partition_covers_global_mesh_eight_ranks(lib_built):
    """Pure host check of the 2x2x2 grid used at 8 GPUs: send lists mirror the peers' halo lists."""
    n, procs = (6, 6, 8), (2, 2, 2)
    parts = [tps_b200.cartesian_hex_partition(n, procs, r, order_mode=1) for r in range(8)]
    gids = np.sort(np.concatenate([p["elem_gid"][:p["num_elems"]] for p in parts]))
    assert np.array_equal(gids, np.arange(n[0] * n[1] * n[2]))
    for r, p in enumerate(parts):
        for pi, qk in enumerate(p["nbr_rank"]):
            sent = p["elem_gid"][p["send_elems"][p["send_offset"][pi]:p["send_offset"][pi + 1]]]
            pq = parts[qk]
            qi = list(pq["nbr_rank"]).index(r)
            recv = pq["elem_gid"][pq["num_elems"] + pq["recv_offset"][qi]:pq["num_elems"] + pq["recv_offset"][qi + 1]]
            assert np.array_equal(sent, recv)


@pytest.mark.parametrize("method", ["metis", "rcb"])
@pytest.mark.parametrize("nparts", [2, 4, 8])
def test_general_partition_of_unstructured_mesh(lib_built, method, nparts):
    """METIS k-way / RCB element maps of the O-grid (config C2 restated) with every element relabelled by a random cube
    rotation, split with tpsb_mk_partition_general: the pieces cover the mesh, send lists mirror the peers' halo lists,
    shared faces agree on both sides, boundary attributes follow their faces."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from common import rotate_elements
    from meshref import HEX_FACE_VERT
    g0 = tps_b200.cylinder_ogrid_mesh(5, 16, 3)
    g = rotate_elements(g0)
    # carry the boundary attributes over the renumbering through the faces' vertex sets
    key = {}
    for f in np.nonzero(g0["face_el2"] < 0)[0]:
        key[tuple(sorted(g0["elem_verts"][g0["face_el1"][f], HEX_FACE_VERT[g0["face_inf1"][f] // 64]].tolist()))] = g0["face_attr"][f]
    attr = np.zeros(len(g["face_el1"]), np.int32)
    for f in np.nonzero(g["face_el2"] < 0)[0]:
        attr[f] = key[tuple(sorted(g["elem_verts"][g["face_el1"][f], HEX_FACE_VERT[g["face_inf1"][f] // 64]].tolist()))]
    g["face_attr"] = attr
    elem_rank, cut = tps_b200.partition_elements(g, nparts, method)
    sizes = np.bincount(elem_rank, minlength=nparts)
    assert sizes.min() > 0 and sizes.max() <= 1.1 * sizes.mean() + 2
    cross = int(((g["face_el2"] >= 0) & (elem_rank[g["face_el1"]] != elem_rank[np.maximum(g["face_el2"], 0)])).sum())
    if cut is not None:
        assert cut == cross
    parts = [tps_b200.partition_mesh(g, elem_rank, r) for r in range(nparts)]
    gids = np.sort(np.concatenate([p["elem_gid"][:p["num_elems"]] for p in parts]))
    assert np.array_equal(gids, np.arange(g["elem_xyz"].shape[0]))
    shared = 0
    for r, p in enumerate(parts):
        ne = p["num_elems"]
        assert (p["face_el1"] < ne).all()
        assert np.array_equal(p["elem_xyz"][:ne], g["elem_xyz"][p["elem_gid"][:ne]])
        shared += int((p["face_el2"] >= ne).sum())
        b = p["face_el2"] < 0
        assert np.array_equal(p["face_attr"][b], g["face_attr"][p["face_gface"][b]]) and (p["face_attr"][b] > 0).all()
        assert (p["face_attr"][~b] == 0).all()
        for pi, qk in enumerate(p["nbr_rank"]):
            sent = p["elem_gid"][p["send_elems"][p["send_offset"][pi]:p["send_offset"][pi + 1]]]
            pq = parts[qk]
            qi = list(pq["nbr_rank"]).index(r)
            recv = pq["elem_gid"][pq["num_elems"] + pq["recv_offset"][qi]:pq["num_elems"] + pq["recv_offset"][qi + 1]]
            assert np.array_equal(sent, recv)
    assert shared == 2 * cross
    local = sum(int(((p["face_el2"] >= 0) & (p["face_el2"] < p["num_elems"])).sum()) for p in parts)
    bdr = sum(int((p["face_el2"] < 0).sum()) for p in parts)
    assert local + cross + bdr == len(g["face_el1"])
