"""Randomised differential test: random chains, filter shapes, unit sizes and strides through the fused
kernels, the general executor and the CPU oracle must agree bit for bit (EXACT mode)."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import assert_bit_equal, gpu_chain, kept_only, oracle_chain, synth_raw

pytestmark = pytest.mark.gpu


def _random_case(seed):
    rng = np.random.default_rng(seed)
    fmt = int(rng.choice([O.CF32, O.CS8, O.CU8, O.CS16]))
    rate = int(rng.choice([48_000, 2_400_000, 20_000_000, 100_000_000]))
    stages = []
    for _ in range(int(rng.integers(0, 3))):
        stages.append(("shift", int(rng.integers(-rate // 2 + 1, rate // 2 - 1))))
    n_lp = int(rng.choice([0, 1, 1, 1, 2]))
    cur_rate = rate
    for _ in range(n_lp):
        D = int(rng.choice([2, 4, 8, 16, 32, 3, 5]))
        L = int(rng.choice([2, 4, 6, 8, 10, 16, 24, 40, 40, 64, 100, 200, int(rng.integers(1, 60)) * 2]))
        stages.append(("lowpass", int(rng.integers(1, max(2, cur_rate // 4))), D, L))
        cur_rate //= D
        if rng.random() < 0.2 and cur_rate > 4:
            stages.append(("shift", int(rng.integers(-cur_rate // 2 + 1, cur_rate // 2 - 1))))
    sink = str(rng.choice(["write", "spark", "spark", "read"]))
    return fmt, rate, stages, sink, rng


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("QD_RANDOM_SEEDS", "96"))))
def test_random_chain(seed):
    import quadrs_b200 as Q

    fmt, rate, stages, sink, rng = _random_case(seed)
    mult = 1
    for st in stages:
        if st[0] == "lowpass":
            mult *= st[2]
    n = int(min(400_000, max(30_000, 600 * mult)))
    raw, _ = synth_raw(fmt, n, seed=0x1000 + seed, rate=rate)
    o = oracle_chain(raw, fmt, rate, stages)
    try:
        o_len = o.len()
    except O.OracleError:
        pytest.skip("capture shorter than the filter")
    g = gpu_chain(raw, fmt, rate, stages)
    ref = gpu_chain(raw, fmt, rate, stages).set_option("use_fast", 0)
    assert g.len() == o_len
    what = f"seed {seed}: fmt {fmt} {stages} {sink}"
    with kept_only():
        if sink == "write":
            chunk = int(rng.choice([64, 512, 0x1000]))
            want, wrc = o.write_mem(chunk=chunk)
            got, grc = g.write_mem(chunk=chunk)
            gen, _ = ref.write_mem(chunk=chunk)
            assert grc == wrc, what
            assert_bit_equal(got, want, what)
            assert_bit_equal(gen, want, what + " (general executor)")
        elif sink == "read":
            for _ in range(4):
                off = int(rng.integers(0, max(1, o_len - 1)))
                cnt = int(rng.integers(1, 3000))
                try:
                    want = o.read_at(off, cnt)
                except O.OracleError as e:
                    with pytest.raises(Q.QdError) as ge:
                        g.read_at(off, cnt)
                    assert ge.value.code == e.code, what
                    continue
                assert_bit_equal(g.read_at(off, cnt), want, f"{what} read_at({off},{cnt})")
        else:
            W = int(rng.choice([1, 2, 4, 8, 16, 32, 64, 128, 256, 1024]))
            if o_len <= W + 2:
                pytest.skip("too few samples for this width")
            S = int(rng.choice([1, 2, max(1, W // 4), W, W + 3]))
            rows = o.spark_rows(W, S)
            max_rows = min(rows, 400)
            first = int(rng.integers(0, rows - max_rows + 1))
            scale = float(rng.choice([1.0, 100.0, 3e4])) if fmt in (O.CU8, O.CS16) else 1.0
            rng_ = (0.02 * scale * np.sqrt(W), 3.0 * scale * np.sqrt(W))
            try:
                widx, wmag = o.spark_fft(W, S, rng_, first_row=first, max_rows=max_rows)
            except O.OracleError as e:
                with pytest.raises(Q.QdError) as ge:
                    g.spark_fft(W, S, rng_, first_row=first, max_rows=max_rows)
                assert ge.value.code == e.code, what
                return
            idx, mag = g.spark_fft(W, S, rng_, first_row=first, max_rows=max_rows, want_mag=True)
            idx2, _ = g.spark_fft(W, S, rng_, first_row=first, max_rows=max_rows, want_mag=False)
            gidx, _ = ref.spark_fft(W, S, rng_, first_row=first, max_rows=max_rows)
            assert np.array_equal(idx, widx), what
            assert np.array_equal(idx2, widx), what + " (threshold epilogue)"
            assert np.array_equal(gidx, widx), what + " (general executor)"
            assert_bit_equal(mag, wmag, what)
