"""Writes the inputs of the reference-pinning fixture (tests/golden/ref_inputs/*.csv) — plain CSV with round-trip
float formatting, readable by Julia's DelimitedFiles without any package.

    python tests/golden/make_reference_inputs.py

The matching outputs come from the REAL BulkLMM.jl: `julia tests/golden/make_reference_fixtures.jl` (needs a Julia
with BulkLMM.jl installed — not available in the build image) writes tests/golden/ref_outputs/*.csv, and
tests/test_reference_fixtures.py then checks both the CPU oracle and the CUDA engine against them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200"))
from blmm_b200 import synth  # noqa: E402

N, P, M = 79, 150, 16


def main():
    out = os.path.join(HERE, "ref_inputs")
    os.makedirs(out, exist_ok=True)
    G = synth.make_geno(N, P, seed=2024)
    K = np.round(synth.calc_kinship_host(G), 12)  # as test/generate_test_bxdData.jl:14 does
    Y = synth.make_pheno(G, K, M, seed=2025)
    Z = synth.make_covar(N, seed=7)
    w = np.random.default_rng(11).uniform(0.5, 1.5, N)
    for name, a in (("G", G), ("K", K), ("Y", Y), ("Covar", Z), ("weights", w.reshape(-1, 1))):
        np.savetxt(os.path.join(out, name + ".csv"), a, delimiter=",", fmt="%.17g")
    print("wrote", out)


if __name__ == "__main__":
    main()
