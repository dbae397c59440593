"""The formulation behind fk_tcfir (rows of 64 samples, frequency-translated taps split in two f16 halves, row phasors,
sum over the rows that meet in an output) as a numpy model against the oracle: catches geometry and indexing errors
without a GPU.  The kernel itself is tested in test_gpu_tcfir.py."""
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from helpers import kept_only, oracle_chain, rel_err, synth_raw

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "scripts"))
import tcfir_model as M  # noqa: E402


@pytest.mark.parametrize("D,L,f", [(8, 40, 1_500_000), (16, 100, 3_000_000), (4, 24, -2_000_000), (32, 40, 700_000),
                                   (2, 18, 1_000_000), (8, 64, None), (8, 38, 9_999_999)])
def test_row_formulation_matches_the_oracle(D, L, f):
    raw, _ = synth_raw(O.CS8, 40_000, rate=20e6)
    st = ([("shift", f)] if f is not None else []) + [("lowpass", 1_000_000, D, L)]
    with kept_only():
        want = oracle_chain(raw, O.CS8, 20_000_000, st).read_at(0, 1000)
    T = (L - L // 2 + D - 1) // D - 1  # the read's truncated tail is not part of the stream the kernel computes
    got = M.model(np.frombuffer(raw, np.int8), 20_000_000, f, 1_000_000, D, L, 0, 1000 - T)
    assert rel_err(got, want[: 1000 - T]) <= 2e-6


def test_geometry_of_the_bench_shape():
    g = M.geometry(40, 8)
    assert (g["OPR"], g["NOUT"], g["N"], g["DMAX"], g["c0"], g["cown"]) == (8, 13, 64, 2, -7, -2)
