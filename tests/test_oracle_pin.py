"""Pins the CPU oracle against every known answer the reference holds for the hot path.

The reference has no hot-path tests; the only quantitative known answer is the
README OOK walkthrough (README.md:112-187) on examples/cupboard-superdec.sr400.cf32, plus the
qualitative config-1 picture (README.md:90-94, screenshots/fsk-5.png).
"""
import itertools
import re

import numpy as np
import pytest

import oracle_lib as O


@pytest.fixture(scope="module")
def cupboard(golden_dir):
    return O.Samples.from_file(golden_dir / "cupboard-superdec.sr400.cf32", O.CF32, 400)


def _bits(idx):
    return "".join("." if (r == 0).all() else "X" for r in idx)


def test_readme_ook_run_lengths(cupboard):
    # README.md:113-116: sparkfft -width 4 -stride 2 -range 0.001:0.01
    assert cupboard.len() == 1994 and cupboard.sample_rate() == 400
    idx, _ = cupboard.spark_fft(4, 2, (0.001, 0.01))
    assert idx.shape == (995, 4)
    runs = [(k, len(list(g))) for k, g in itertools.groupby(_bits(idx))]
    # README.md:135-140: " 8 . / 8 X / 16 . / 17 X / 15 . / 16 X"
    want = [(".", 8), ("X", 8), (".", 16), ("X", 17), (".", 15), ("X", 16)]
    assert any(runs[i : i + 6] == want for i in range(len(runs))), runs


def test_readme_ook_bitstrings_and_temperature(cupboard):
    idx, _ = cupboard.spark_fft(4, 2, (0.001, 0.01))
    # the README pipes the header line through sed too, which becomes the leading X
    bits = "X" + _bits(idx).replace(".", "o")
    # README.md:160: sed -E 's/X{6,10}/A/g; s/o{5,10}/B/g'
    ab = re.sub(r"o{5,10}", "B", re.sub(r"X{6,10}", "A", bits))
    assert ab == ("XBBBBBBBBBBBBBBBBBBBBBBBBBBBBBABABABABABABBABAABABABBABAABABABABBAABABBABAABABBAABBAABABABABABABBAABB"
                  "ABBBBBBBBBBBBBooo")  # README.md:161
    # README.md:167-168
    pairs = re.sub(r"(..)", r"\1_", re.sub(r".*BBBBABAB(AB)*BABA", "", ab))
    assert pairs == ("AB_AB_AB_BA_BA_AB_AB_AB_AB_BA_AB_AB_BA_BA_AB_AB_BA_AB_BA_AB_AB_AB_AB_AB_AB_BA_AB_BA_"
                     "BB_BB_BB_BB_BB_BB_Bo_oo_")
    # README.md:174-175
    out = re.sub(r"(.{8})(.)", r"\1^\2^", pairs.replace("AB_", "0").replace("BA_", "1"))
    assert out.startswith("00011000^0^10011001^0^10000001^0^1")
    assert 24 + 153 / 255 == pytest.approx(24.6, abs=1e-9)  # README.md:187


def test_readme_text_format(cupboard):
    txt = cupboard.spark_fft_text(4, 2, (0.001, 0.01))
    lines = txt.split("\n")
    assert lines[0] == "sparkfft sample_rate=400"  # fft.rs:19
    assert len(lines) == 1 + 995 + 1 and lines[-1] == ""
    assert lines[1] == "│    │"  # fft.rs:63
    assert all(len(l) == 6 for l in lines[1:-1])


def test_config1_fsk_structure(golden_dir):
    # README.md:90-94 = BASELINE.json configs[0]
    s = O.Samples.from_file(golden_dir / "fsk-example.sr21M.fc32", O.CF32, 21_000_000)
    assert s.len() == 196_864
    chain = s.shift(280_000).lowpass(200_000, decimate=32, size=400)
    assert chain.len() == 1 + (196_864 - 400) // 32 == 6140
    assert chain.sample_rate() == 21_000_000 // 32
    O.set_kept_only(True)
    try:
        idx, mag = chain.spark_fft(64, 16)
    finally:
        O.set_kept_only(False)
    assert idx.shape == (380, 64)
    assert idx.max() <= 3 and mag.max() < 0.25
    # two alternating FSK columns (screenshots/fsk-5.png): energy concentrates in two bin groups
    col = (idx > 0).sum(axis=0)
    hot = set(np.argsort(col)[-4:].tolist())
    assert hot == {24, 25, 47, 48}, (hot, col)
    # the two tones alternate: rows where the left group is lit rarely have the right group lit
    left = (idx[:, 24:26] > 0).any(axis=1)
    right = (idx[:, 47:49] > 0).any(axis=1)
    assert (left ^ right).mean() > 0.8


def test_config1_against_the_readme_screenshot(golden_dir):
    """screenshots/fsk-5.png (README.md:90-94) is a bilevel image of the reference's own terminal output for config 1:
    45 text rows of 19 px, 64 glyph cells of 10 px (first cell at x = 3), a white bar where the glyph is not blank.
    The picture is periodic (the FSK alternates every ~5 rows), so the row offset is found by best agreement; at it
    99.5 % of the 2880 cells agree with the oracle (2865), and every difference is a single cell at the edge of one of
    the two FSK columns -- magnitudes next to the 0.08 threshold -- so this pins the pipeline's structure and scale
    (which bins, which rows, the threshold) against a run of the real reference, not bit-exactness."""
    Image = pytest.importorskip("PIL.Image")
    a = np.array(Image.open(golden_dir / "readme-fsk-5.png").convert("L")) > 127
    assert a.shape == (854, 640)
    shot = np.zeros((45, 64), dtype=bool)
    for r in range(45):
        band = a[12 + 19 * r - 8: 12 + 19 * r + 6]
        for c in range(63):
            shot[r, c] = band[:, 3 + 10 * c: 13 + 10 * c].any()
    assert set(np.nonzero(shot.any(axis=0))[0].tolist()) == {24, 25, 47, 48}
    c = O.Samples.from_file(golden_dir / "fsk-example.sr21M.fc32", O.CF32, 21_000_000).shift(280_000).lowpass(200_000, 32, 400)
    O.set_kept_only(True)
    try:
        idx, _ = c.spark_fft(64, 16)
    finally:
        O.set_kept_only(False)
    agree = np.array([((idx[r0:r0 + 45] > 0) == shot).sum() for r0 in range(idx.shape[0] - 44)])
    assert agree.max() >= 0.99 * shot.size, agree.max()
    r0 = int(agree.argmax())
    bad = np.argwhere((idx[r0:r0 + 45] > 0) != shot)
    assert all(int(col) in (24, 25, 47, 48) for _, col in bad)  # only the edges of the two FSK columns differ


def test_oracle_reproduces_committed_golden_vectors(golden_dir):
    # tests/golden/*.npy were produced by tests/golden/make_golden.py with the literal convolve
    s = O.Samples.from_file(golden_dir / "cupboard-superdec.sr400.cf32", O.CF32, 400)
    idx, mag = s.spark_fft(4, 2, (0.001, 0.01))
    assert np.array_equal(idx, np.load(golden_dir / "cupboard_idx.npy"))
    assert np.array_equal(mag.view(np.uint32), np.load(golden_dir / "cupboard_mag.npy").view(np.uint32))
    c = O.Samples.from_file(golden_dir / "fsk-example.sr21M.fc32", O.CF32, 21_000_000)
    c = c.shift(280_000).lowpass(200_000, 32, 400)
    O.set_kept_only(True)
    try:
        idx, mag = c.spark_fft(64, 16)
    finally:
        O.set_kept_only(False)
    assert np.array_equal(idx, np.load(golden_dir / "config1_idx.npy"))
    assert np.array_equal(mag.view(np.uint32), np.load(golden_dir / "config1_mag.npy").view(np.uint32))
