"""CPU-side checks: the C-ABI library builds, loads, exports every declared symbol, fails loudly
without a GPU, and its host-only arithmetic (shard planning) matches the reference's counts."""
import re
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def Q():
    import quadrs_b200

    quadrs_b200.build()
    return quadrs_b200


def test_header_symbols_all_exported(Q):
    hdr = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "quadrs_gpu.h").read_text(), flags=re.S)
    declared = set(re.findall(r"\b(qd_[a-z0-9_]+)\s*\(", hdr))
    lib = Q._lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(Q._lib.ABI_SYMBOLS), declared ^ set(Q._lib.ABI_SYMBOLS)
    assert lib.qd_abi_version() == 2


def test_status_codes_shared_with_oracle(Q):
    L = Q._lib
    for name in ("E_INVALID_ARG", "E_SHIFT_NYQUIST", "E_ZERO_RATE", "E_OFFSET_EOF", "E_SHORT_INPUT", "E_SHORT_READ",
                 "E_FFT_WIDTH", "E_GLYPH_RANGE", "E_LEVELS", "E_SLICE", "E_VISIBLE", "E_GEN_ARGS", "E_WRITE_SHORT",
                 "E_IO", "E_UNIMPLEMENTED", "E_EXISTS", "E_NOMEM", "E_ZERO_STRIDE"):
        assert getattr(L, name) == getattr(O, name), name
    assert L.status_name(L.E_WRITE_SHORT) == "QD_E_WRITE_SHORT"


def test_no_cpu_fallback(Q):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Q.QdError) as e:
        Q.Samples.from_bytes(np.zeros(64, dtype=np.uint8), Q.CS8, 1000)
    assert e.value.code == Q._lib.E_CUDA and "no CPU fallback" in e.value.msg


def test_construction_errors_precede_device_use(Q):
    # Shift::new / Gen::new / LowPass argument checks are host logic and must match the oracle's codes
    with pytest.raises(Q.QdError) as e:
        Q.Samples.from_bytes(np.zeros(64, dtype=np.uint8), Q.CS8, 1000).shift(500)
    assert e.value.code in (Q._lib.E_SHIFT_NYQUIST, Q._lib.E_CUDA)
    import ctypes as C
    src = Q._lib.Source()
    src.kind, src.sample_rate, src.gen_seconds, src.gen_n_cos = Q._lib.SRC_GEN, 1000, 1.0, 0
    h = C.c_void_p()
    assert Q._lib.lib().qd_chain_create(C.byref(src), None, 0, 0, C.byref(h)) == Q._lib.E_GEN_ARGS
    st = (Q._lib.Stage * 1)()
    st[0].kind, st[0].frequency = Q._lib.STAGE_SHIFT, 600
    src = Q._lib.Source()
    src.kind, src.format, src.sample_rate, src.n_bytes = Q._lib.SRC_HOST_MEM, Q.CS8, 1000, 0
    assert Q._lib.lib().qd_chain_create(C.byref(src), st, 1, 0, C.byref(h)) == Q._lib.E_SHIFT_NYQUIST
    assert b"half the sample rate" in Q._lib.lib().qd_last_error()


def test_format_row_matches_oracle(Q):
    idx = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 0], dtype=np.uint8)
    assert Q.format_row(idx) == O.format_row(idx) == "│ ▁▂▃▄▅▆▇█ │"


@pytest.mark.parametrize("cfg", [
    # (fmt, total, stages, sink, unit, stride)
    (O.CS8, 2**30, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], 0, 4096, 4096),
    (O.CS16, 2**33, [("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)], 1, 128, 128),
    (O.CF32, 2**34, [("lowpass", 20_000_000, 8, 40), ("lowpass", 500_000, 32, 40)], 1, 4, 2),
    (O.CU8, 2**30, [], 1, 4096, 1024),
    (O.CF32, 196_864, [("shift", 280_000), ("lowpass", 200_000, 32, 400)], 1, 64, 16),
])
@pytest.mark.parametrize("n_shards", [1, 2, 8])
def test_shard_plan_covers_units_once_with_halo(Q, cfg, n_shards):
    fmt, total, stages, sink, unit, stride = cfg
    rate = 100_000_000
    plans = Q.plan_shards(fmt, rate, total, stages, sink, unit, stride, n_shards)
    # units: contiguous, disjoint, complete
    assert plans[0].first_unit == 0
    for a, b in zip(plans, plans[1:]):
        assert a.first_unit + a.n_units == b.first_unit
    ln = total
    mult, need = 1, unit
    for st in reversed(stages):
        if st[0] == "lowpass":
            need = need * st[2] + st[3]
            mult *= st[2]
    for st in stages:
        if st[0] == "lowpass":
            ln = 1 + (ln - st[3]) // st[2]
    n_units = -(-ln // unit) if sink == 0 else (-(-(ln - unit) // stride) if ln > unit else 0)
    assert plans[-1].first_unit + plans[-1].n_units == n_units
    step = unit if sink == 0 else stride
    for p in plans:
        if p.n_units == 0:
            continue
        assert p.first_sample == p.first_unit * step * mult
        last_end = min((p.first_unit + p.n_units - 1) * step * mult + need, total)
        assert p.first_sample + p.n_samples == last_end
    # halo between adjacent shards = (unit*D + L) - step*D samples
    if n_shards > 1 and plans[0].n_units and plans[1].n_units:
        overlap = plans[0].first_sample + plans[0].n_samples - plans[1].first_sample
        assert overlap == need - step * mult


def test_sharded_chain_argument_checks_need_no_device(Q):
    """qd_chain_create_sharded validates its arguments before it touches CUDA (include/quadrs_gpu.h)."""
    import ctypes as C
    L = Q._lib
    lib = L.lib()
    src = L.Source()
    src.kind, src.format, src.sample_rate = L.SRC_HOST_MEM, L.FMT_CS8, 1000
    buf = np.zeros(64, dtype=np.uint8)
    src.data, src.n_bytes = buf.ctypes.data, buf.size
    h = C.c_void_p()
    devs = (C.c_int * 2)(0, 1)
    assert lib.qd_chain_create_sharded(C.byref(src), None, 0, None, 2, C.byref(h)) == L.E_INVALID_ARG
    assert lib.qd_chain_create_sharded(C.byref(src), None, 0, devs, 0, C.byref(h)) == L.E_INVALID_ARG
    src.kind = L.SRC_DEVICE_MEM  # a device-resident capture lives on one device
    assert lib.qd_chain_create_sharded(C.byref(src), None, 0, devs, 2, C.byref(h)) == L.E_INVALID_ARG
    assert b"one device" in lib.qd_last_error()
    assert h.value is None
