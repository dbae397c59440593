"""fk_tcfir (qd_tcfir.cu): the FAST-arithmetic filter on the tensor cores (tcgen05.mma, accumulators in TMEM) for cs8
captures, against the oracle (north_star tolerance: cf32 samples within 1e-5 relative, max-norm per 0x1000-sample
chunk), against the CUDA-core FAST kernel, and bit for bit against itself across host segments, chunk sizes and
capture windows."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import gpu_chain, kept_only, oracle_chain, rel_err, synth_raw

pytestmark = pytest.mark.gpu

TC_NAME = "fk_tcfir"


@pytest.fixture(scope="module")
def Q():
    import quadrs_b200

    return quadrs_b200


def _ran_on_tensor_cores(chain, fn):
    chain.profile(True)
    out = fn()
    _, _, name = chain.profile_read()
    chain.profile(False)
    return out, TC_NAME in name


SHAPES = [
    # rate, stages, base sample, chunk
    (20_000_000, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], 0, 0x1000),  # config 2
    (20_000_000, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], 2000 * 8 * 0x1000, 0x1000),                                                    # deep into a capture
    (20_000_000, [("lowpass", 1_000_000, 8, 40)], 0, 0x1000),                        # no shift
    (20_000_000, [("shift", -3_000_000), ("lowpass", 500_000, 16, 100)], 0, 0x1000),
    (20_000_000, [("shift", 2_000_000), ("lowpass", 3_000_000, 4, 24)], 0, 0x1000),
    (20_000_000, [("shift", 700_000), ("lowpass", 300_000, 32, 40)], 0, 512),
    (20_000_000, [("shift", 1_000_000), ("shift", -2_500_000), ("lowpass", 1_000_000, 8, 64)], 0, 1000),
    (2_400_000, [("shift", 100_000), ("lowpass", 100_000, 8, 100)], 0, 0x1000),
]


def _mult(stages):
    m = 1
    for st in stages:
        if st[0] == "lowpass":
            m *= st[2]
    return m


@pytest.mark.parametrize("rate,stages,base,chunk", SHAPES)
def test_tensor_core_fir_within_1e5_of_the_oracle(Q, rate, stages, base, chunk):
    D = _mult(stages)
    n = chunk * D * 12 + 5000
    total = base + n
    raw, _ = synth_raw(O.CS8, n, first=base, rate=rate)
    first = base // (chunk * D)
    assert first * chunk * D == base
    with kept_only():
        want, _ = oracle_chain(raw, O.CS8, rate, stages, base, total if base else 0).write_mem(chunk=chunk, first_chunk=first, max_chunks=10)
    tc = gpu_chain(raw, O.CS8, rate, stages, base, total if base else 0, precision=Q.FAST)
    (got, _), ran = _ran_on_tensor_cores(tc, lambda: tc.write_mem(chunk=chunk, first_chunk=first, max_chunks=10))
    assert ran, "the tensor-core kernel did not run for this shape"
    assert len(got) == len(want) == 10 * chunk
    worst = max(rel_err(got[c : c + chunk], want[c : c + chunk]) for c in range(0, len(want), chunk))
    print(f"fk_tcfir worst chunk rel err {worst:.3e}")
    assert worst <= 1e-5, worst
    # ... and the CUDA-core FAST kernel agrees to the same tolerance (two independent FAST paths)
    cc = gpu_chain(raw, O.CS8, rate, stages, base, total if base else 0, precision=Q.FAST).set_option("use_tc", 0)
    (ref, _), ran_cc = _ran_on_tensor_cores(cc, lambda: cc.write_mem(chunk=chunk, first_chunk=first, max_chunks=10))
    assert not ran_cc
    assert max(rel_err(got[c : c + chunk], ref[c : c + chunk]) for c in range(0, len(want), chunk)) <= 1e-5


def test_shapes_the_tensor_core_kernel_declines_run_on_the_cuda_cores(Q):
    """A long filter whose exchange buffers exceed shared memory, and n * ratio >= 2^29 (the reference's phase rounding,
    which only the CUDA-core kernel re-applies, would show)."""
    rate = 20_000_000
    for stages, base in (([("shift", 1_000_000), ("lowpass", 1_000_000, 8, 400)], 0),
                         ([("shift", 1_000_000), ("lowpass", 4_000_000, 2, 18)], 0),
                         ([("shift", 9_999_999), ("lowpass", 1_000_000, 8, 40)], 2**30 // (8 * 0x1000) * 8 * 0x1000)):
        D = _mult(stages)
        n = 0x1000 * D * 4 + 5000
        raw, _ = synth_raw(O.CS8, n, first=base, rate=rate)
        first = base // (0x1000 * D)
        with kept_only():
            want, _ = oracle_chain(raw, O.CS8, rate, stages, base, base + n if base else 0).write_mem(first_chunk=first, max_chunks=3)
        h = gpu_chain(raw, O.CS8, rate, stages, base, base + n if base else 0, precision=Q.FAST)
        (got, _), ran = _ran_on_tensor_cores(h, lambda: h.write_mem(first_chunk=first, max_chunks=3))
        assert not ran
        assert max(rel_err(got[c : c + 0x1000], want[c : c + 0x1000]) for c in range(0, len(want), 0x1000)) <= 1e-5


def test_truncated_tails_are_the_exact_arithmetic(Q):
    """The T outputs at the end of every read use a truncated filter (filter.rs:68-71): fk_tail writes them in the exact
    arithmetic, so they equal the oracle bit for bit also in FAST mode."""
    rate, stages, chunk = 20_000_000, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], 256
    raw, _ = synth_raw(O.CS8, chunk * 8 * 9 + 500, rate=rate)
    with kept_only():
        want, _ = oracle_chain(raw, O.CS8, rate, stages).write_mem(chunk=chunk, max_chunks=8)
    tc = gpu_chain(raw, O.CS8, rate, stages, precision=Q.FAST)
    (got, _), ran = _ran_on_tensor_cores(tc, lambda: tc.write_mem(chunk=chunk, max_chunks=8))
    assert ran
    T = (20 + 7) // 8 - 1
    w, g = want.reshape(8, chunk), got.reshape(8, chunk)
    assert np.array_equal(w[:, chunk - T :].view(np.uint32), g[:, chunk - T :].view(np.uint32))
    assert not np.array_equal(w.view(np.uint32), g.view(np.uint32))  # the body is FAST arithmetic


def test_results_do_not_depend_on_segments_windows_or_read_sizes(Q):
    """Rows sit at absolute sample positions and every output sums its rows in a fixed order, so the FAST output is
    bit-identical whatever the host-path segment size, the resident window of the capture or the read that asks."""
    rate, stages, chunk = 20_000_000, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], 0x1000
    n = chunk * 8 * 40 + 3000
    raw, _ = synth_raw(O.CS8, n, rate=rate)
    base_chain = gpu_chain(raw, O.CS8, rate, stages, precision=Q.FAST)
    (want, _), ran = _ran_on_tensor_cores(base_chain, lambda: base_chain.write_mem(chunk=chunk, max_chunks=40))
    assert ran and len(want) == 40 * chunk
    for seg in (100_000, 777_777, 3_000_000):
        h = gpu_chain(raw, O.CS8, rate, stages, precision=Q.FAST).set_option("segment_bytes", seg)
        got, _ = h.write_mem(chunk=chunk, max_chunks=40)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), seg
    # a window of the capture that starts mid-row (absolute sample 5 * 8 * 0x1000 - 24): same outputs for the chunks it holds
    lo = 5 * 8 * chunk
    sub = raw[2 * (lo - 24) :]
    h = gpu_chain(sub, O.CS8, rate, stages, lo - 24, n, precision=Q.FAST)
    got, _ = h.write_mem(chunk=chunk, first_chunk=5, max_chunks=20)
    assert np.array_equal(got.view(np.uint32), want[5 * chunk : 25 * chunk].view(np.uint32))
    # one read in the middle of a chunk: the untruncated part equals the chunked output
    T = (20 + 7) // 8 - 1
    part = base_chain.read_at(3 * chunk + 17, 1000)
    assert np.array_equal(part[: 1000 - T].view(np.uint32), want[3 * chunk + 17 : 3 * chunk + 17 + 1000 - T].view(np.uint32))


def test_device_resident_capture_and_output(Q):
    import torch

    rate, n, chunk = 20_000_000, 2**24, 0x1000
    synth = Q.make_synth(0x5EED0002, [(Q.tone_step(1.6e6, rate), 45, 0), (Q.tone_step(-4.1e6, rate), 30, 0)], 6)
    d_in = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    Q.synth_fill_device(synth, Q.CS8, 0, n, d_in.data_ptr())
    torch.cuda.synchronize()
    chunks = n // 8 // chunk - 1
    outs = []
    for prec, tc in ((Q.EXACT, 1), (Q.FAST, 1), (Q.FAST, 0)):
        chain = Q.Samples.from_device(d_in.data_ptr(), 2 * n, Q.CS8, rate, keep=(d_in,)).shift(1_500_000).lowpass(1_000_000, 8, 40)
        chain = chain.with_precision(prec).set_option("use_tc", tc)
        d_out = torch.zeros(2 * chunks * chunk, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()  # the fill runs on torch's stream, the chain on its own (non-blocking) stream
        got_n, rc = chain.write_into(chunk, 0, chunks, d_out.data_ptr(), chunks * chunk, Q._lib.SPACE_DEVICE)
        chain.synchronize()
        assert got_n == chunks * chunk and rc == 0
        outs.append(torch.view_as_complex(d_out.view(-1, 2)).view(chunks, chunk))
    exact, tcf, ccf = outs
    err = (tcf - exact).abs().amax(dim=1) / exact.abs().amax(dim=1)
    print(f"fk_tcfir vs EXACT over 2^24 samples: worst chunk rel err {float(err.max()):.3e}, mean {float(err.mean()):.3e}")
    assert float(err.max()) <= 1e-5
    err2 = (ccf - exact).abs().amax(dim=1) / exact.abs().amax(dim=1)
    print(f"CUDA-core FAST vs EXACT: worst {float(err2.max()):.3e}")
