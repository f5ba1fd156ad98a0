#!/usr/bin/env python
"""A/B comparison of two saved events, the role of the reference's tests/compare_ab.py:

    python scripts/compare_ab.py <A event dir or record.npy> <B event dir or record.npy> [--unshifted] [--atol 1e-5]

prints the photon indices whose step records differ, and - when both sides hold seq.npy - the history table with chi2
(ana/qcf.py role)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _load(path, name):
    p = path if path.endswith(".npy") else os.path.join(path, name)
    return np.load(p) if os.path.exists(p) else None


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("a"); ap.add_argument("b")
    ap.add_argument("--unshifted", action="store_true", help="compare step k with step k (two GPU events) instead of A[k+1] with B[k]")
    ap.add_argument("--atol", type=float, default=1e-5)
    a = ap.parse_args(argv)
    from eic_opticks_b200 import analysis as A
    ra, rb = _load(a.a, "record.npy"), _load(a.b, "record.npy")
    print(ra.shape); print(rb.shape)
    diff = A.compare_ab(ra, rb, atol=a.atol, shifted=not a.unshifted)
    print(diff)
    sa, sb = (None, None) if a.a.endswith(".npy") else (_load(a.a, "seq.npy"), _load(a.b, "seq.npy"))
    if sa is not None and sb is not None:
        c2, ndf, rows = A.chi2_histories(sa, sb)
        print("chi2/ndf %.2f / %d" % (c2, ndf))
        for r in rows[:20]:
            print(r)
    return 0


if __name__ == "__main__":
    sys.exit(main())
