"""CPU-side evidence from the shipped SASS (cuobjdump of the in-tree build; no GPU needed).

* the fused kernel stages tiles with the bulk-copy engine (UBLKCP + mbarrier SYNCS) and computes in packed FP32
  (FFMA2 / FMUL2), as DESIGN.md says;
* the long-filter loop of fk_fir takes its taps from the constant bank through UNIFORM registers (LDCU) and not
  through per-thread constant loads (LDC): ptxas decides that from the shape of the code, and has silently fallen
  back to LDC several times while the loop was being written (6 % slower on config 4), so it is pinned here.
"""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "quadrs_b200" / "_build"


@pytest.fixture(scope="module")
def built():
    import quadrs_b200

    quadrs_b200.build()
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    return BUILD


def functions(obj: Path):
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
    for f in out.split("Function : ")[1:]:
        name = f.split("\n")[0].strip()
        ins, addr = [], []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
            if m:
                addr.append(int(m.group(1), 16))
                ins.append(m.group(2))
        yield name, addr, ins


def loops(addr, ins):
    pos = {a: i for i, a in enumerate(addr)}
    for i, t in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr[i] and tgt in pos:
                yield ins[pos[tgt]: i + 1]


@pytest.mark.parametrize("obj,kernel", [
    ("qd_fast_d16.o", "fk_firILi16ELi4ELi128ELb1ELi0ELb0E"),   # config 4: cs16, L = 800, EXACT
    ("qd_fast_d16.o", "fk_firILi16ELi4ELi128ELb1ELi0ELb1E"),   # the same with window-tail snapshots
    ("qd_fast_d32.o", "fk_firILi32ELi2ELi128ELb1ELi0ELb1E"),   # config 1: cf32, L = 400, EXACT, snapshots
    ("qd_fast_d32.o", "fk_firILi32ELi2ELi128ELb1ELi0ELb0E"),
])
def test_long_filter_loop_loads_taps_through_uniform_registers(built, obj, kernel):
    found = False
    for name, addr, ins in functions(built / obj):
        if kernel not in name:
            continue
        found = True
        steady = [b for b in loops(addr, ins) if 100 <= sum("FMUL2" in x or "FFMA2" in x for x in b) < 1000]
        assert steady, "no steady-state filter loop found"
        body = max(steady, key=lambda b: sum("FMUL2" in x or "FFMA2" in x for x in b))
        math = sum("FMUL2" in x or "FFMA2" in x for x in body)
        ldcu = sum(x.startswith("LDCU") for x in body)
        ldc = sum(re.match(r"(@!?P\d\s+)?LDC[.\s]", x) is not None for x in body)
        assert ldcu >= 16 and ldc == 0, f"{name}: steady loop has {ldcu} LDCU and {ldc} LDC for {math} packed math instructions"
        assert math / len(body) >= 0.6, f"{name}: {math} math of {len(body)} instructions"
    assert found, kernel


def test_bulk_copy_and_packed_fp32_are_in_the_shipped_kernels(built):
    text = subprocess.run(["cuobjdump", "-sass", str(built / "qd_fast_d8.o")], capture_output=True, text=True, check=True).stdout
    for mnemonic in ("UBLKCP", "SYNCS", "FFMA2", "FMUL2"):
        assert mnemonic in text, mnemonic


def test_tensor_core_fir_uses_tcgen05_tmem_and_the_bulk_copy_engine(built):
    """fk_tcfir (qd_tcfir.cu): UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld (accumulators read back from TMEM),
    UTCBAR = tcgen05.commit onto an mbarrier, UBLKCP = the 16 KB bulk copies of raw bytes; and no I2F in the int8 -> f16
    conversion (PRMT into the mantissa of 1024.0h, one packed subtraction: no I2F to f16)."""
    for name, addr, ins in functions(built / "qd_tcfir.o"):
        if "fk_tcfir" not in name:
            continue
        text = "\n".join(ins)
        for mnemonic in ("UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "PRMT", "HADD2"):
            assert mnemonic in text, (name, mnemonic)
        assert sum(x.startswith("UTCHMMA") or " UTCHMMA" in x for x in ins) >= 8, name  # K = 128 as eight K = 16 steps
        assert not any(re.search(r"I2F(P)?\.F16", x) for x in ins), name  # int8 -> f16 without conversion instructions


def test_stft_glyphs_through_the_linear_form_and_no_twiddle_negations(built):
    """fk_stft<12> (config 3): the epilogue decides glyphs from MUFU.SQRT + NaN-propagating clamps (qd_stft_epilogue.cuh
    glyph_lin), and no FADD builds (-wy, wx) for the twiddle products any more (pmul_tw: the sign rides on the
    constant operand of the FFMA2) -- both were FMA-pipe / issue costs that bounded the kernel."""
    for name, _, ins in functions(built / "qd_stft.o"):
        if "fk_stftILi12ELi4E" not in name:
            continue
        ops = [t.split()[1] if t.startswith("@") else t.split()[0] for t in ins]
        assert sum(o.startswith("MUFU.SQRT") for o in ops) >= 16, "one square root per bin of a thread"
        assert sum(o.startswith("FMNMX.NAN") for o in ops) >= 32
        assert not any(re.match(r"FADD R\d+, -R\d+, -RZ", re.sub(r"^@!?P\d\s+", "", t)) for t in ins), "explicit negations are back"
        return
    raise AssertionError("fk_stft<12, 4> not found in qd_stft.o")


def test_stft_padded_offsets_are_linear():
    import importlib.util

    spec = importlib.util.spec_from_file_location("check_stft_pad", ROOT / "scripts" / "check_stft_pad.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.check() == 0
