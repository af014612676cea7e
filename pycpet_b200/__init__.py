"""pycpet_b200 -- B200 (sm_100a) implementation of PyCPET's data-parallel hot path.

Coulomb field / ESP of M point charges at N points (`volume`, `volume_ESP`), the fixed-step
streamline tracer with fused distance + curvature (`topo` / `topo_GPU`) and the
distance x curvature histogram, behind PyCPET's own C-shared-library pattern:

* ``pycpet_b200.c_ops.Math_ops``        mirror of CPET/utils/c_ops.py (ctypes over libcpetb200.so)
* ``pycpet_b200.calculator``            mirror of the calculator-level entry points
* ``pycpet_b200.device.Engine``         device-pointer API (torch tensors in/out, no host copies)
* ``pycpet_b200.sharding``              one-process-per-GPU partitioning + NCCL gather/all-reduce
* ``pycpet_b200.md_batch``              MD-trajectory driver: PyCPET's constructor in worker processes,
                                        pipelined with the batched GPU call
* ``pycpet_b200.io``                    `.top` / `.dat` writers (np.savetxt's bytes) and readers
* ``include/cpet_b200.h``               the C ABI itself

There is no CPU / PyTorch fallback: without the built CUDA library and an sm_100 device every
compute call raises ``CpetError``.
"""
from ._lib import CpetError, lib_path
from .c_ops import Math_ops
from .calculator import (
    compute_ESP_on_grid,
    compute_box,
    compute_box_ESP,
    compute_field_on_grid,
    compute_point_field,
    compute_topo_GPU_batch_filter,
    compute_topo_batch,
    compute_topo_complete_c_shared,
    construct_distance_matrix,
    distance_numpy,
    get_math,
    make_histograms,
    patch_reference,
)

__all__ = [
    "CpetError", "Math_ops", "lib_path", "get_math", "patch_reference",
    "compute_field_on_grid", "compute_ESP_on_grid", "compute_topo_batch",
    "compute_box", "compute_box_ESP", "compute_point_field",
    "compute_topo_complete_c_shared", "compute_topo_GPU_batch_filter",
    "make_histograms", "distance_numpy", "construct_distance_matrix",
]
__version__ = "0.1.0"
