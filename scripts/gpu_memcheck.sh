#!/bin/bash
# compute-sanitizer memcheck over a slice of the GPU tests that covers every kernel variant (small inputs)
set -u
mkdir -p gpurun_out
timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest -x -q -m gpu \
  "tests/test_gpu_fast.py::test_overlapping_windows_stream_plus_tail" \
  "tests/test_gpu_fast.py::test_fast_output_is_independent_of_segments_and_shards" \
  "tests/test_gpu_fast.py::test_fast_sparkfft_is_independent_of_segments" \
  "tests/test_gpu_fast.py::test_sharded_equals_unsharded" \
  "tests/test_gpu_fast.py::test_shards_and_pointers_at_awkward_alignments" \
  "tests/test_gpu_multi.py" \
  "tests/test_gpu_parity.py::test_read_at_bit_exact_including_truncated_tail" \
  "tests/test_gpu_parity.py::test_sparkfft_bucket_indices_bit_exact" \
  "tests/test_gpu_tcfir.py" \
  > gpurun_out/r2_memcheck.log 2>&1
echo "exit $?" >> gpurun_out/r2_memcheck.log
tail -15 gpurun_out/r2_memcheck.log
