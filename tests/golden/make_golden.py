"""Writes tests/golden/vren_tcnn_small.npz from the oracle (run in this container: `python tests/golden/make_golden.py`).
The reference ships no golden vectors and its kernels cannot be built or imported here (F1/F2), so
these fixtures freeze the oracle's own outputs on small seeded inputs; tests/test_oracle.py::test_golden_fixture_matches
recomputes and compares them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from test_oracle import _golden_cases  # noqa: E402

if __name__ == "__main__":
    cases = _golden_cases()
    np.savez_compressed(os.path.join(HERE, "vren_tcnn_small.npz"), **{k: v.numpy() for k, v in cases.items()})
    print({k: tuple(v.shape) for k, v in cases.items()})
