"""Regenerates the golden vectors under tests/golden/.

Inputs: the two capture fixtures of the reference (examples/cupboard-superdec.sr400.cf32 and
examples/fsk-example.sr21M.fc32, copied here byte for byte as DATA fixtures -- /root/reference does
not exist on the GPU box).  Outputs: bucket-index matrices produced by the CPU oracle
(oracle/quadrs_oracle.c).  They are oracle outputs, NOT outputs of a real quadrs binary: no Rust
toolchain exists in the build image.  What pins them to the reference is the README known answer
(tests/test_oracle_pin.py), which the cupboard matrix reproduces.

Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_lib as O  # noqa: E402


def main():
    s = O.Samples.from_file(HERE / "cupboard-superdec.sr400.cf32", O.CF32, 400)
    idx, mag = s.spark_fft(4, 2, (0.001, 0.01))
    np.save(HERE / "cupboard_idx.npy", idx)
    np.save(HERE / "cupboard_mag.npy", mag)
    c = O.Samples.from_file(HERE / "fsk-example.sr21M.fc32", O.CF32, 21_000_000).shift(280_000).lowpass(200_000, 32, 400)
    idx, mag = c.spark_fft(64, 16)  # literal complex_convolve, not the kept-only shortcut
    np.save(HERE / "config1_idx.npy", idx)
    np.save(HERE / "config1_mag.npy", mag)
    print("cupboard", np.load(HERE / "cupboard_idx.npy").shape, "config1", idx.shape)


if __name__ == "__main__":
    main()
