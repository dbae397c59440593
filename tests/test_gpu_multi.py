"""Single-process multi-GPU executor (qd_chain_create_sharded): one host process, one sink, several devices.

The reference's caller is one process folding commands into one sink (src/bin/quadrs.rs:48-56); the sharded
chain fans every sink call out over the devices by contiguous unit ranges and gathers the results into the
caller's one buffer.  On a one-GPU box the same code runs with one device named several times (every shard its
own chain, streams, staging and host thread); with more GPUs visible the real devices are used as well."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import assert_bit_equal, gpu_chain, kept_only, oracle_chain, synth_raw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Q():
    import quadrs_b200

    quadrs_b200.build()
    return quadrs_b200


def device_sets(Q):
    import torch

    n = torch.cuda.device_count()
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets.append(list(range(min(n, 8))))
    return sets


CASES = [
    (O.CS16, 100_000_000, [("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)], (128, 128, (0.5, 50.0))),
    (O.CS8, 20_000_000, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], (64, 16, (0.01, 3.0))),
    (O.CF32, 400_000_000, [("lowpass", 20_000_000, 8, 40), ("lowpass", 500_000, 32, 40)], (4, 2, (0.001, 0.01))),
    (O.CU8, 2_400_000, [], (1024, 256, (2.0, 500.0))),
]


@pytest.mark.parametrize("fmt,rate,stages,spark", CASES)
@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_sharded_chain_equals_one_device_chain(Q, fmt, rate, stages, spark, precision):
    if precision == "fast" and fmt in (O.CU8, O.CS16):
        pytest.skip("FAST is refused for the offset formats")
    prec = Q.FAST if precision == "fast" else Q.EXACT
    n = 900_000
    raw, _ = synth_raw(fmt, n, rate=rate)
    one = gpu_chain(raw, fmt, rate, stages, precision=prec)
    W, S, rng = spark
    want_idx, want_mag = one.spark_fft(W, S, rng, want_mag=True)
    want_w, want_rc = one.write_mem()
    want_lv, want_total = one.freq_levels(W, S)
    for devs in device_sets(Q):
        many = one.on_devices(devs).with_precision(prec)
        assert many.n_devices() == len(devs)
        assert many.len() == one.len() and many.sample_rate() == one.sample_rate()
        many.set_option("segment_bytes", 150_000)  # several pipelined segments per device
        idx, mag = many.spark_fft(W, S, rng, want_mag=True)
        assert np.array_equal(idx, want_idx), devs
        assert_bit_equal(mag, want_mag, f"sparkfft magnitudes on {devs}")
        got_w, rc = many.write_mem()
        assert rc == want_rc
        assert_bit_equal(got_w, want_w, f"write on {devs}")
        lv, total = many.freq_levels(W, S)
        assert total == want_total and np.array_equal(lv, want_lv)
        # a sub-range of rows, as a pager would ask for
        sub, _ = many.spark_fft(W, S, rng, first_row=7, max_rows=101)
        assert np.array_equal(sub, want_idx[7:108])


def test_sharded_chain_matches_oracle_and_file_source(Q, tmp_path):
    n = 400_000
    raw, _ = synth_raw(O.CS8, n)
    st = [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)]
    with kept_only():
        want_idx, _ = oracle_chain(raw, O.CS8, 20_000_000, st).spark_fft(64, 16, (0.01, 3.0))
        want_w, want_rc = oracle_chain(raw, O.CS8, 20_000_000, st).write_mem()
    path = tmp_path / "cap.sr20M.cs8"
    raw.tofile(path)
    f = Q.Samples.from_file(path, Q.CS8, 20_000_000, device=[0, 0, 0]).shift(1_500_000).lowpass(1_000_000, 8, 40)
    idx, _ = f.spark_fft(64, 16, (0.01, 3.0))
    assert np.array_equal(idx, want_idx)
    got, rc = f.write_mem()
    assert rc == want_rc
    assert_bit_equal(got, want_w, "sharded file source, write")
    name = f.write_file(str(tmp_path / "out"), overwrite=True)
    on_disk = np.fromfile(name, dtype=np.complex64)
    assert_bit_equal(on_disk, want_w, "sharded write_file")
    # take_fft rows are spread over the devices too
    with kept_only():
        want_t = oracle_chain(raw, O.CS8, 20_000_000, st).take_fft(256, 300, None, True)
    got_t = f.take_fft(256, 300, None, True)
    assert_bit_equal(got_t, want_t, "sharded take_fft")


def test_sharded_chain_errors(Q):
    import torch

    raw, _ = synth_raw(O.CS8, 50_000)
    d = torch.from_numpy(raw).cuda()
    with pytest.raises(Q.QdError) as e:
        Q.Samples.from_device(d.data_ptr(), raw.size, Q.CS8, 20_000_000, device=[0, 0])
    assert e.value.code == Q._lib.E_INVALID_ARG
    many = gpu_chain(raw, O.CS8, 20_000_000, [("lowpass", 1_000_000, 8, 40)]).on_devices([0, 0])
    # the end-of-capture Err of the last window and the write panic surface exactly as on one device
    one = gpu_chain(raw, O.CS8, 20_000_000, [("lowpass", 1_000_000, 8, 40)])
    a, rc_a = one.write_mem()
    b, rc_b = many.write_mem()
    assert rc_a == rc_b == Q._lib.E_WRITE_SHORT
    assert_bit_equal(b, a, "ragged end")
    with pytest.raises(Q.QdError) as e:
        many.with_stream(0)
    assert e.value.code == Q._lib.E_INVALID_ARG


def test_large_stride_windows_stream_in_small_segments(Q):
    """Windows much further apart than wide (an overview of a long capture) with a small staging budget: the
    segments are sized from the real span (stride per window), so the call streams instead of failing."""
    n = 3_000_000
    raw, _ = synth_raw(O.CS8, n)
    for stages in ([("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], [("lowpass", 1_000_000, 8, 16)], []):
        W, S = 128, 20_000 if stages else 160_000
        with kept_only():
            want, _ = oracle_chain(raw, O.CS8, 20_000_000, stages).spark_fft(W, S, (0.01, 3.0))
        g = gpu_chain(raw, O.CS8, 20_000_000, stages)
        g.set_option("segment_bytes", 1 << 20)
        g.set_option("scratch_budget", 4 << 20)
        got, _ = g.spark_fft(W, S, (0.01, 3.0))
        assert np.array_equal(got, want), stages
