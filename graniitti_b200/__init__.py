"""graniitti_b200 -- B200-native batched event-weight engine for GRANIITTI's hot path.

The product is the CUDA library graniitti_b200/libgrb.so behind the C ABI of include/grb.h; this package
is the thin ctypes plumbing used by the tests and bench.py.  There is no CPU fallback: importing works
without a GPU (so that host-side parsing can be tested), every compute call needs one."""
from .engine import (Engine, GrbError, lib, parse_card, optimize_grid, measure_dfma_peak, local_calls, comm_unique_id, event_record, write_events, writer_opts,  # noqa: F401
                     eikonal_cache_names, write_eikonal_file, read_eikonal_file, test_fastmath)
from .dist import allreduce_accum, DevArray  # noqa: F401
from . import _abi  # noqa: F401
