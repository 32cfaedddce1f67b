"""B200-native orbital-optimisation inner loop (energy, analytic gradient, Stiefel retraction,
Barzilai-Borwein step) behind the reference's PartialUnitaryProjectionOptimizer API.

The numerical work is done by hand-written sm_100a CUDA kernels in liboo_b200.so (C ABI:
include/oo_b200.h); this package is the host-side mirror of the reference interface."""
from . import _lib, distributed, ingest, synthetic
from .engine import OrbitalEngine, measure_peaks
from .optimizer import PartialUnitaryProjectionOptimizer, clear_engine_cache, content_checksum
from .distributed import shard_range, attach_nccl, attach_peer_memory
from .ingest import SpatialIntegrals
from .rotated import RotatedHamiltonianMixin, patch_reference, rotated_spin_integrals

__all__ = ["PartialUnitaryProjectionOptimizer", "OrbitalEngine", "measure_peaks", "shard_range",
           "attach_nccl", "attach_peer_memory", "clear_engine_cache", "content_checksum",
           "SpatialIntegrals", "RotatedHamiltonianMixin", "patch_reference",
           "rotated_spin_integrals", "ingest", "synthetic"]
