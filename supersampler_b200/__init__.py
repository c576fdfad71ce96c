"""supersampler_b200 -- B200-native sketch-and-compare hot path of SuperSampler.

The product is native code: `lib/libspsp_b200.so` (hand-written sm_100a CUDA
kernels behind the C ABI of include/spsp.h) and `lib/libspsp_host.so` (the C++
host layer mirroring the reference's Subsampler / Comparator classes), plus the
drop-in executables `bin/sub_sampler` and `bin/comparator`.  This Python
package only builds them and binds them with ctypes for tests and bench.py.
There is no CPU fallback: GPU entry points raise when no CUDA device exists.
"""
from .capi import (  # noqa: F401
    SpspError, build, device_lib, host_lib, threshold, pack_fasta, postpass, decode_sketch,
    format_csv, write_csv_gz, sketch_buffers, compare_buffers, run_sub_sampler, run_comparator, run_sort_csv, DeviceContext,
    HIT_DTYPE, packed_words, batch_layout, Sketcher, Comparer, Pipeline, BatchStream, PinnedBuffer, postpass_batch, SCAN_AUTO, SCAN_DENSE, SCAN_FILTER,
)
