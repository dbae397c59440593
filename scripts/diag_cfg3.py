import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import quadrs_b200 as Q, oracle_lib as O
from test_gpu_fast import _oracle_rows_from_device
rate, total, W, S, rng = 2_400_000, 2**30, 4096, 1024, (2.0, 500.0)
synth = Q.make_synth(0x5EED0003, [(Q.tone_step(-800e3, rate), 40, 0), (Q.tone_step(-123_456, rate), 30, 0), (Q.tone_step(300e3, rate), 25, 0), (Q.tone_step(1_000_001, rate), 20, 0)], 4)
d_in = torch.empty(2 * total + 64, dtype=torch.uint8, device="cuda")
Q.synth_fill_device(synth, Q.CU8, 0, total, d_in.data_ptr()); torch.cuda.synchronize()
whole = Q.Samples.from_device(d_in.data_ptr(), 2 * total, Q.CU8, rate, keep=(d_in,))
rows_total = whole.spark_rows(W, S)
d_idx = torch.zeros(rows_total * W, dtype=torch.uint8, device="cuda")
whole.spark_fft_device(W, S, rng, 0, rows_total, d_idx.data_ptr()); whole.synchronize()
rows = [0, 1, 1000, 262143, 524285, 524286, 524287, 524288, 524289, 700000, rows_total - 2, rows_total - 1]
want = _oracle_rows_from_device(Q, torch, d_in, Q.CU8, rate, total, 0, [], W, S, rng, rows, 4096)
for r, (widx, wmag) in zip(rows, want):
    got = d_idx[r * W:(r + 1) * W].cpu().numpy()
    bad = np.nonzero(got != widx)[0]
    print("row", r, "differing bins", len(bad), bad[:6], got[bad[:6]], widx[bad[:6]], (wmag[bad[:6]] if wmag is not None else None))
# the same rows computed on their own (small call)
for r in (524287, 524288):
    d2 = torch.zeros(W, dtype=torch.uint8, device="cuda")
    whole.spark_fft_device(W, S, rng, r, 1, d2.data_ptr()); whole.synchronize()
    print("row", r, "single-row call equals whole-run row:", bool(torch.equal(d2, d_idx[r * W:(r + 1) * W])))
