"""poutine_b200 — B200-native (sm_100a) homoplasy counting + permutation null for POUTINE.

Only the hot path of the reference (`Homoplasy_Counter.count_all_homoplasy_events` + `assoc`) lives
here; see DESIGN.md.  Everything numeric runs in libpoutine_b200.so (hand-written CUDA, C ABI in
include/poutine_b200.h); this package is the thin host mirror used by tests and bench.
"""
from ._lib import LIB_PATH, build  # noqa: F401
from .counter import HomoplasyCounter, HomoplasyCounterGroup, NoInformativeSites, PoutineError, comm_unique_id, compact_planes, pack_chars, pack_chars_biallelic  # noqa: F401
