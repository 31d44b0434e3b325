"""Loader of the golden vectors: tests/golden/hotpath_golden.npz (inputs + oracle-made outputs, tests/golden/make_golden.py),
overlaid with tests/golden/hotpath_golden_julia.npz when a maintainer has produced it with the real package
(tests/golden/make_golden.jl).  `provenance` tells which outputs the parity tests were held to."""
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load():
    g = dict(np.load(os.path.join(HERE, "hotpath_golden.npz")))
    jl = os.path.join(HERE, "hotpath_golden_julia.npz")
    provenance = "oracle (NumPy restatement; parity unpinned for raw reference outputs)"
    if os.path.exists(jl):
        j = np.load(jl)
        for k in j.files:
            g[k] = j[k]
        provenance = "julia (TensorTrainNumerics.jl outputs, make_golden.jl) for: " + ", ".join(sorted(j.files))
    g["__provenance__"] = provenance
    return g
