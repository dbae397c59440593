"""FAST arithmetic over a device-resident cs8 capture: tensor-core kernel against the CUDA-core kernel, per filter shape
(device time of the chain's kernels by CUDA events, qd_chain_profile)."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import quadrs_b200 as Q

rate, n, chunk = 20_000_000, 2**28, 0x1000
synth = Q.make_synth(0x5EED0002, [(Q.tone_step(1.6e6, rate), 45, 0), (Q.tone_step(-4.1e6, rate), 30, 0)], 6)
d_in = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
Q.synth_fill_device(synth, Q.CS8, 0, n, d_in.data_ptr())
torch.cuda.synchronize()
for D, L in [(8, 40), (8, 64), (8, 24), (16, 40), (16, 100), (4, 24), (4, 40), (32, 40), (32, 100), (2, 18), (8, 100)]:
    chunks = n // D // chunk - 2
    d_out = torch.zeros(2 * chunks * chunk, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    row = []
    for tc in (1, 0):
        c = Q.Samples.from_device(d_in.data_ptr(), 2 * n, Q.CS8, rate, keep=(d_in,)).shift(1_500_000).lowpass(1_000_000, D, L)
        c = c.with_precision(Q.FAST).set_option("use_tc", tc)
        for _ in range(2):
            c.write_into(chunk, 0, chunks, d_out.data_ptr(), chunks * chunk, Q._lib.SPACE_DEVICE)
        c.synchronize()
        c.profile(True)
        for _ in range(3):
            c.write_into(chunk, 0, chunks, d_out.data_ptr(), chunks * chunk, Q._lib.SPACE_DEVICE)
        _, ms, name = c.profile_read()
        c.profile(False)
        row.append((n * 3 / (ms * 1e-3) / 1e9, name.split(" ")[0]))
    print(f"D={D:2d} L={L:3d}  tc: {row[0][0]:7.1f} Gs/s ({row[0][1]})   cuda cores: {row[1][0]:7.1f} Gs/s ({row[1][1]})  ratio {row[0][0] / row[1][0]:.2f}")
