"""host/quadrs_gpu: the command-line front end with the reference's grammar (src/args.rs, src/bin/quadrs.rs)."""
import subprocess
from pathlib import Path

import pytest

import oracle_lib as O

ROOT = Path(__file__).resolve().parent.parent
CLI = ROOT / "host" / "quadrs_gpu"


@pytest.fixture(scope="module")
def cli():
    import quadrs_b200

    quadrs_b200.build()
    subprocess.run(["make", "-C", str(ROOT / "host")], check=True, capture_output=True)
    return str(CLI)


def run(cli, *args):
    return subprocess.run([cli, *args], capture_output=True, text=True)


def test_readme_command_line_parses_to_the_reference_operations(cli):
    r = run(cli, "--parse-only", "from", "fsk-example.sr21M.fc32", "shift", "280000", "lowpass", "-power", "200",
            "-decimate", "32", "200000", "sparkfft", "-width", "64", "-stride", "16")
    assert r.returncode == 0
    assert r.stdout.splitlines() == [
        "From filename=fsk-example.sr21M.fc32 format=cf32 sample_rate=21000000",  # args.rs:328-333,392-402
        "Shift frequency=280000",
        "LowPass size=400 decimate=32 frequency=200000",  # -power P -> size 2P (args.rs:161-166)
        "SparkFft width=64 stride=16",
    ]


def test_defaults_suffixes_and_sniffing(cli):
    r = run(cli, "--parse-only", "from", "gqrx_20180126_111922_868000000_8000000_fc.raw", "shift", "-47k", "lowpass", "2M",
            "sparkfft", "-range", "0.001:0.01", "bucket", "-by", "freq", "2", "write", "-overwrite", "yes", "out",
            "gen", "-cos", "1k", "-cos", "-2500", "-len", "0.5", "48k")
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines() == [
        "From filename=gqrx_20180126_111922_868000000_8000000_fc.raw format=cf32 sample_rate=8000000",  # args.rs:111-118
        "Shift frequency=-47000",                       # "-47k": third char is a digit -> a number (args.rs:421-426)
        "LowPass size=40 decimate=8 frequency=2000000",  # defaults args.rs:165,170
        "SparkFft width=128 stride=128 min=0.00100000005 max=0.00999999978",  # stride defaults to width (args.rs:195-198)
        "Bucket fft_width=128 stride=128 levels=2",
        "Write overwrite=true prefix=out",
        "Gen sample_rate=48000 seconds=0.5 cos=1000,-2500",  # -cos may repeat (args.rs:35)
    ]
    r = run(cli, "--parse-only", "from", "-sr", "2400k", "-format", "cu8", "capture.bin", "from", "g001_433.92M_250k.cu8",
            "from", "x.sr1G.c16")
    assert r.stdout.splitlines() == [
        "From filename=capture.bin format=cu8 sample_rate=2400000",
        "From filename=g001_433.92M_250k.cu8 format=cu8 sample_rate=250000",  # rtl_433 pattern args.rs:120-125
        "From filename=x.sr1G.c16 format=cs16 sample_rate=1000000000",
    ]


@pytest.mark.parametrize("args,needle", [
    ([], "no commands provided"),
    (["frobnicate"], "unrecognised command"),
    (["from", "nosuch.bin"], "unable to guess sample rate"),
    (["from", "x.sr1M.xyz"], "unable to guess format"),
    (["shift"], "'shift' requires a frequency argument"),
    (["lowpass", "-power", "3", "-power", "4", "1000"], "specified more than once"),
    (["lowpass", "-bogus", "1", "1000"], "invalid flags"),
    (["sparkfft", "-range", "1"], "range argument must contain a ':'"),
    (["bucket", "-by", "time", "2"], "must bucket -by freq"),
    (["gen", "48k"], "gen requires at least one operation"),
    (["write", "-overwrite", "maybe", "x"], "unacceptable boolean value"),
])
def test_errors_print_usage_then_the_message(cli, args, needle):
    r = run(cli, "--parse-only", *args)
    assert r.returncode == 1
    assert r.stdout.startswith("usage: ") and " lowpass [-power 20] [-decimate 8] FREQUENCY \\" in r.stdout  # quadrs.rs:9-28
    assert needle in r.stderr


@pytest.mark.gpu
def test_readme_ook_example_stdout_is_byte_identical(cli, golden_dir):
    fixture = golden_dir / "cupboard-superdec.sr400.cf32"
    r = run(cli, "from", str(fixture), "sparkfft", "-width", "4", "-stride", "2", "-range", "0.001:0.01")
    assert r.returncode == 0, r.stderr
    want = O.Samples.from_file(fixture, O.CF32, 400).spark_fft_text(4, 2, (0.001, 0.01))
    assert r.stdout == want
    assert r.stdout.splitlines()[0] == "sparkfft sample_rate=400"


@pytest.mark.gpu
def test_config1_pipeline_and_write_roundtrip(cli, golden_dir, tmp_path):
    import numpy as np

    fixture = golden_dir / "fsk-example.sr21M.fc32"
    r = run(cli, "from", str(fixture), "shift", "280000", "lowpass", "-power", "200", "-decimate", "32", "200000",
            "sparkfft", "-width", "64", "-stride", "16")
    assert r.returncode == 0, r.stderr
    chain = O.Samples.from_file(fixture, O.CF32, 21_000_000).shift(280_000).lowpass(200_000, 32, 400)
    O.set_kept_only(True)
    try:
        assert r.stdout == chain.spark_fft_text(64, 16)
        want, _ = chain.write_mem()
    finally:
        O.set_kept_only(False)
    prefix = tmp_path / "dec"
    r = run(cli, "from", str(fixture), "shift", "280000", "lowpass", "-power", "200", "-decimate", "32", "200000",
            "write", str(prefix))
    assert r.returncode == 101 and "short read at offset" in r.stderr  # lib.rs:203 panics after the data is out
    data = np.fromfile(str(prefix) + ".sr656250.cf32", dtype=np.complex64)
    assert np.array_equal(data.view(np.uint32), want.view(np.uint32))
    r = run(cli, "gen", "-cos", "1000", "-len", "0.01", "48k", "bucket", "-width", "64", "-by", "freq", "2")
    assert r.returncode == 0 and set(r.stdout.strip()) <= {"0", "1"} and len(r.stdout.strip()) == (480 - 64) // 64


@pytest.mark.gpu
def test_gpus_option_shards_one_process_over_the_devices(cli, golden_dir, tmp_path):
    """`quadrs_gpu --gpus N ...`: the same command line on N devices of the box (qd_chain_create_sharded): stdout and
    the written file are byte-identical to the one-device run."""
    import numpy as np
    import torch

    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("one GPU visible")
    fixture = golden_dir / "fsk-example.sr21M.fc32"
    cmd = ["from", str(fixture), "shift", "280000", "lowpass", "-power", "200", "-decimate", "32", "200000"]
    one = run(cli, *cmd, "sparkfft", "-width", "64", "-stride", "16")
    many = run(cli, "--gpus", str(n), *cmd, "sparkfft", "-width", "64", "-stride", "16")
    assert one.returncode == 0 and many.returncode == 0, many.stderr
    assert many.stdout == one.stdout
    r1 = run(cli, *cmd, "write", str(tmp_path / "a"))
    rn = run(cli, "--gpus", str(n), *cmd, "write", str(tmp_path / "b"))
    assert r1.returncode == rn.returncode == 101
    a = np.fromfile(str(tmp_path / "a") + ".sr656250.cf32", dtype=np.uint32)
    b = np.fromfile(str(tmp_path / "b") + ".sr656250.cf32", dtype=np.uint32)
    assert np.array_equal(a, b)
