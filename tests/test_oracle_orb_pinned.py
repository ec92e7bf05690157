"""Pins the ORB restatement (oracle/extras.py) and the product's host state machine (ipplb_orb_begin / next / cut / finish)
against the REAL OrthogonalRecursiveBisection of the reference -- findCutAxis, findMedian, cutDomain,
perpendicularReduction, binaryRepartition and FieldLayout::updateLayout executing the reference's own code on serial
stand-ins (oracle/ref_shim/refshim_orb.cpp) -- live where /root/reference exists and everywhere through the committed
tests/golden/ref_orb.npz (tests/golden/make_golden_orb.py).  Boxes are integer: exact."""
import os
import sys

import numpy as np
import pytest

import ippl_b200 as ib
from oracle import extras as ox
from oracle import refshim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_orb import GRIDS, MEDIANS, RANKS, weights  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_orb.npz"))


def _state_machine(ng, nranks, w):
    orb = ib.Orb(ng, nranks)
    while True:
        nxt = orb.next()
        if nxt is None:
            break
        lo, hi, axis = nxt
        sub = w[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
        orb.cut(sub.sum(axis=tuple(a for a in range(3) if a != 2 - axis)))
    return orb.finish()


@pytest.mark.parametrize("gi", range(len(GRIDS)))
@pytest.mark.parametrize("kind", ["uniform", "blob", "random"])
def test_repartition_vs_reference(gold, gi, kind):
    ng = GRIDS[gi]
    w = gold[f"w_{gi}_random"] if kind == "random" else weights(ng, kind, 100 + gi)
    for nr in RANKS:
        want, want_ok = gold[f"boxes_{gi}_{kind}_{nr}"], bool(gold[f"ok_{gi}_{kind}_{nr}"][0])
        ob, ook = ox.orb_repartition(ng, nr, w)                     # the numpy restatement
        pb, pok = _state_machine(ng, nr, w)                         # the product's C-ABI state machine
        assert ook == want_ok == pok, (ng, kind, nr)
        if want_ok:   # a rejected repartition leaves the reference's layout untouched: nothing to compare
            assert np.array_equal(np.asarray(ob, dtype=np.int32), want), (ng, kind, nr)
            assert np.array_equal(pb, want), (ng, kind, nr)
        if refshim.orb_available():
            lb, lok = refshim.orb_repartition(ng, nr, w)
            assert lok == want_ok and (not lok or np.array_equal(lb, want))


def test_find_median_vs_reference(gold):
    for i, w in enumerate(MEDIANS):
        want = int(gold[f"median_{i}"][0])
        assert ox.orb_find_median([float(x) for x in w]) == want, w
        if refshim.orb_available():
            assert refshim.orb_find_median(w) == want
        if len(w) >= 3:   # through the product: one cut of a (len, 2, 2) domain along x
            orb = ib.Orb((len(w), 2, 2), 2)
            orb.next()
            orb.cut(np.asarray(w, dtype=np.float64))
            boxes, _ = orb.finish()
            assert boxes[0][3] == want and boxes[1][0] == want + 1


def test_scatter_loop_vs_reference(gold):
    """scatterR executed by the reference's code: l = (R - origin) * invdx + 0.5, index = (int) l, whi = l - index,
    wlo = 1 - whi, args = index - lDom.first() + nghost, scatterToField -- the particle loop ParticleAttrib::scatter
    shares (ParticleAttrib.hpp:167-184).  The serial restatement visits particles and stencil points in the same order:
    bit-exact, corner / face / centre particles included."""
    import oracle
    ng, origin, h = tuple(int(v) for v in gold["sr_ng"]), tuple(gold["sr_origin"]), tuple(gold["sr_h"])
    R = [gold["sr_x"], gold["sr_y"], gold["sr_z"]]
    m = oracle.Mesh.make(ng, origin, h)
    got = oracle.field_zeros(m)
    oracle.scatter_cic(m, *R, 1.0, got)
    assert np.array_equal(got, gold["sr_field"])
    assert abs(got.sum() - len(R[0])) < 1e-9
    if refshim.orb_available():
        assert np.array_equal(refshim.orb_scatter_r(ng, origin, h, *R), gold["sr_field"])


def test_neighbour_tables_of_orb_layouts_vs_reference(gold):
    """After an ORB repartition the reference rebuilds its halo tables with FieldLayout::updateLayout -> findNeighbors
    (FieldLayout.hpp:59-73, 203-341) on boxes of unequal size.  The restatement and the product's ipplb_layout_set_boxes
    + ipplb_layout_neighbors reproduce components, peers and send / receive ranges exactly."""
    import oracle
    n = 0
    for gi, ng in enumerate(GRIDS):
        for kind in ("blob", "random"):
            for nr in (2, 3, 4, 5, 8):
                if f"nb_{gi}_{kind}_{nr}_0" not in gold:
                    continue
                boxes = gold[f"boxes_{gi}_{kind}_{nr}"]
                L = ib.Layout(ng, nr)
                L.set_boxes(boxes)
                for my in range(nr):
                    want = gold[f"nb_{gi}_{kind}_{nr}_{my}"]
                    assert np.array_equal(oracle.neighbors(ng, boxes, my), want), (ng, kind, nr, my)
                    assert np.array_equal(L.neighbors(my), want), (ng, kind, nr, my)
                    if refshim.available():
                        assert np.array_equal(refshim.neighbors_boxes(ng, boxes, my), want)
                    n += 1
                L.close()
    assert n >= 100
