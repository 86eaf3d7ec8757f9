"""Host logic of the chunked host-buffer pipeline (tpsb_rhs_mult_host): the static schedule must respect every data
dependency of the evaluation -- a chunk's gradient after the primitives of all its face neighbours' chunks, a face range
after the gradients (trace blocks) of both sides' chunks, a chunk's residual after all its faces -- for any chunk count,
including the periodic wrap (chunk 0 needs the last chunk).  No device needed."""
import numpy as np
import pytest

import tps_b200


def _check(m, chunks):
    sched = tps_b200.host_pipe_schedule(m, chunks)
    assert sched is not None
    eb, fb, ops = sched
    NE, NF = m["elem_xyz"].shape[0], len(m["face_el1"])
    assert eb[0] == 0 and eb[-1] == NE and fb[0] == 0 and fb[-1] == NF
    assert (np.diff(eb) > 0).all() and (np.diff(fb) >= 0).all()
    chunk_of = lambda e: int(np.searchsorted(eb, e, side="right") - 1)
    fchunk_of = lambda f: int(np.searchsorted(fb, f, side="right") - 1)
    el1, el2 = m["face_el1"], m["face_el2"]
    # every face sits in the face range of Elem1's chunk
    for f in range(NF):
        assert fchunk_of(f) == chunk_of(el1[f])
    # neighbours and faces per element chunk
    nbr_chunks = [set([c]) for c in range(chunks)]
    face_chunks_of_elem_chunk = [set() for _ in range(chunks)]
    face_needs = [set() for _ in range(chunks)]
    for f in range(NF):
        c1, c2, cf = chunk_of(el1[f]), chunk_of(el2[f]), fchunk_of(f)
        nbr_chunks[c1].add(c2), nbr_chunks[c2].add(c1)
        face_chunks_of_elem_chunk[c1].add(cf), face_chunks_of_elem_chunk[c2].add(cf)
        face_needs[cf].update((c1, c2))
    done = {0: set(), 1: set(), 2: set(), 3: set()}
    empty_face_chunks = {c for c in range(chunks) if fb[c + 1] == fb[c]}
    for kind, c in ops:
        if kind == 0:
            assert c == len(done[0])  # copies arrive in chunk order
        elif kind == 1:
            assert nbr_chunks[c] <= done[0], (c, nbr_chunks[c], done[0])
        elif kind == 2:
            assert face_needs[c] <= done[1]
        else:
            assert c in done[1] and face_chunks_of_elem_chunk[c] <= (done[2] | empty_face_chunks)
        assert c not in done[kind]
        done[kind].add(c)
    full = set(range(chunks))
    assert done[0] == full and done[1] == full and done[3] == full and done[2] == full - empty_face_chunks
    return ops


@pytest.mark.parametrize("n,chunks", [((6, 6, 6), 4), ((5, 4, 7), 7), ((8, 8, 8), 32), ((4, 4, 4), 3), ((3, 3, 12), 12), ((4, 4, 4), 64)])
def test_schedule_respects_dependencies(lib_built, n, chunks):
    m = tps_b200.cartesian_hex_mesh(*n)
    _check(m, chunks)


def test_schedule_overlaps_early_chunks_with_later_copies(lib_built):
    """With z-slab chunks only the wrap-around chunks wait for the last copy: most residuals are issued (and their
    copy-out can start) before the last chunk has even arrived."""
    m = tps_b200.cartesian_hex_mesh(4, 4, 16)
    ops = _check(m, 16)
    last_copy = ops.index((0, 15))
    early_resid = [c for k, (kind, c) in enumerate(ops) if kind == 3 and k < last_copy]
    assert len(early_resid) >= 11


def test_unchunkable_inputs_are_refused(lib_built):
    m = tps_b200.cartesian_hex_mesh(3, 3, 3)
    assert tps_b200.host_pipe_schedule(m, 2) is None      # fewer than 3 chunks: not worth a pipeline
    assert tps_b200.host_pipe_schedule(m, 28) is None     # more chunks than elements


@pytest.mark.parametrize("n,periodic,chunks", [((4, 4, 6), (0, 0, 0), 5), ((5, 3, 8), (1, 0, 1), 8), ((3, 3, 3), (0, 0, 0), 3)])
def test_schedule_with_boundary_faces(lib_built, n, periodic, chunks):
    """Boundary faces (BCintegrator) ride in the face range of their element's chunk: every chunk's residual comes after
    the face op of each chunk holding one of its two-sided faces and after its own boundary-face range."""
    m = tps_b200.cartesian_hex_mesh(*n, periodic=periodic)
    sched = tps_b200.host_pipe_schedule(m, chunks, with_bdr=True)
    assert sched is not None
    eb, fb, ops, bb = sched
    el1, el2 = np.asarray(m["face_el1"]), np.asarray(m["face_el2"])
    two = np.flatnonzero(el2 >= 0)
    bdr = np.flatnonzero(el2 < 0)
    assert fb[-1] == len(two) and bb[-1] == len(bdr) and len(bdr) > 0
    chunk_of = lambda e: int(np.searchsorted(eb, e, side="right") - 1)
    for k, f in enumerate(bdr):  # boundary face k sits in the boundary range of its element's chunk
        assert bb[chunk_of(el1[f])] <= k < bb[chunk_of(el1[f]) + 1]
    need = [set() for _ in range(chunks)]
    for i, f in enumerate(two):
        cf = int(np.searchsorted(fb, i, side="right") - 1)
        assert cf == chunk_of(el1[f])
        need[chunk_of(el1[f])].add(cf), need[chunk_of(el2[f])].add(cf)
    for f in bdr:
        need[chunk_of(el1[f])].add(chunk_of(el1[f]))
    faced, graded = set(), set()
    for kind, c in ops:
        if kind == 1:
            graded.add(c)
        elif kind == 2:
            assert c in graded
            faced.add(c)
        elif kind == 3:
            assert need[c] <= faced, (c, need[c], faced)
    assert {c for k, c in ops if k == 3} == set(range(chunks))
