"""oracle/check_golden.py — TEST INFRASTRUCTURE ONLY.

Diff a re-generated golden directory against the committed one:
  UNETCA_GOLDEN_OUT=/tmp/g python oracle/make_golden.py && python oracle/check_golden.py /tmp/g
Every array of every .npz must be bit-identical (the fixtures are seeded and the reference runs single-process fp32 / fp64
on the CPU); exits non-zero otherwise."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    other = sys.argv[1]
    gold = os.path.join(ROOT, "tests", "golden")
    bad = 0
    for f in sorted(os.listdir(other)):
        if not f.endswith(".npz"):
            continue
        a, b = np.load(os.path.join(gold, f)), np.load(os.path.join(other, f))
        keys = sorted(set(a.files) | set(b.files))
        worst = 0.0
        for k in keys:
            if k not in a.files or k not in b.files:
                print(f"{f}: key {k} only on one side"); bad += 1
                continue
            x, y = a[k], b[k]
            if x.shape != y.shape or x.dtype != y.dtype:
                print(f"{f}:{k} shape/dtype differ"); bad += 1
                continue
            if not np.array_equal(x, y, equal_nan=True) if x.dtype.kind == "f" else not np.array_equal(x, y):
                d = float(np.max(np.abs(x.astype(np.float64) - y.astype(np.float64)))) if x.dtype.kind in "fiu" else 1.0
                worst = max(worst, d)
                print(f"{f}:{k} differs, max abs {d:.3e}"); bad += 1
        print(f"{f}: {len(keys)} arrays, {'identical' if worst == 0.0 else 'DIFFERENT'}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
