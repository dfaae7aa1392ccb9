"""Golden vectors for the minimal-representation path, produced by the UNMODIFIED reference compiled into oracle/_ref:
daqp_minrep(is_redundant, A, b, n, m, ms) (reference include/api.h:54, src/api.c:531-556, src/utils.c:808-835) on
seeded polyhedra. The reference's own test-suites hold no minrep case, so these outputs of the reference itself are the
pin. Run in the build container (needs /root/reference):

    python tests/golden/make_golden_minrep.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import harness  # noqa: E402
from daqp_b200.problems import generate_polyhedra  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (P, n, m, ms, seed)
CASES = {
    "minrep_n2_m8": (16, 2, 8, 0, 11),
    "minrep_n3_m20": (16, 3, 20, 0, 12),
    "minrep_n5_m40_ms2": (12, 5, 40, 2, 13),
    "minrep_n10_m100": (8, 10, 100, 0, 14),
    "minrep_n10_m60_ms10": (8, 10, 60, 10, 15),
    "minrep_n20_m150": (4, 20, 150, 0, 16),
    "minrep_n50_m300": (2, 50, 300, 0, 17),
    # every third polyhedron made EMPTY by a contradictory pair of rows (the reference's answer is then the one its
    # probing order produces: constraints are dropped from the front until the rest is non-empty)
    "minrep_n4_m24_empty": (24, 4, 24, 0, 18),
    "minrep_n6_m30_ms3_empty": (18, 6, 30, 3, 19),
}


def make_some_empty(A, b, ms, seed):
    rng = np.random.default_rng(seed)
    mA = A.shape[1]
    for q in range(0, A.shape[0], 3):
        j1, j2 = rng.choice(mA, 2, replace=False)
        A[q, j2] = -A[q, j1]
        b[q, ms + j2] = -b[q, ms + j1] - rng.uniform(0.5, 2.0)
    return A, b


def special_cases():
    """Hand-made polyhedra for the edge cases: duplicated rows, a scaled duplicate, an all-zero row, an empty
    polyhedron, an unbounded one, and the unit box with a cut."""
    out = {}
    box = np.vstack([np.eye(2), -np.eye(2)])
    out["box_with_cuts"] = (np.vstack([box, [[1.0, 1.0]], [[1.0, 1.0]], [[1.0, 0.0]]]),
                            np.array([1.0, 1.0, 1.0, 1.0, 1.5, 3.0, 2.0]))
    out["duplicates"] = (np.vstack([box, box[:1], 2.0 * box[1:2]]), np.array([1.0, 1.0, 1.0, 1.0, 1.0, 2.0]))
    out["zero_row"] = (np.vstack([box, [[0.0, 0.0]]]), np.array([1.0, 1.0, 1.0, 1.0, 0.5]))
    out["empty"] = (np.array([[1.0, 0.0], [-1.0, 0.0], [0.0, 1.0]]), np.array([-1.0, -1.0, 1.0]))
    out["unbounded"] = (np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]), np.array([1.0, 2.0, 1.0, 5.0]))
    out["empty_two_rounds"] = (np.array([[1.0, 0.0], [1.0, 0.0], [-1.0, 0.0], [0.0, 1.0]]), np.array([-1.0, -2.0, -1.0, 1.0]))
    out["single"] = (np.array([[1.0, 2.0, 3.0]]), np.array([1.0]))
    return out


def main():
    harness.build(ref=True)
    for name, (P, n, m, ms, seed) in CASES.items():
        A, b = generate_polyhedra(P, n, m, ms, seed)
        if name.endswith("_empty"):
            A, b = make_some_empty(A, b, ms, seed)
        red = np.stack([harness.ref_minrep(A[q], b[q]) for q in range(P)])
        strict = np.stack([harness.ref_minrep(A[q], b[q], "libdaqp_ref_strict.so") for q in range(P)])
        assert (red == strict).all(), name
        np.savez_compressed(os.path.join(HERE, name + ".npz"), A=A, b=b, n=n, m=m, ms=ms, is_redundant=red)
        print(name, "redundant", int(red.sum()), "of", red.size)
    sp = special_cases()
    pack = {}
    for k, (A, b) in sp.items():
        red = harness.ref_minrep(A, b)
        assert (red == harness.ref_minrep(A, b, "libdaqp_ref_strict.so")).all(), k
        pack[k + "_A"] = A; pack[k + "_b"] = b; pack[k + "_red"] = red
        print(k, red.tolist())
    np.savez_compressed(os.path.join(HERE, "minrep_special.npz"), names=np.array(sorted(sp)), **pack)


if __name__ == "__main__":
    main()
