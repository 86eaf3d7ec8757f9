"""tps_b200 -- B200-native (sm_100a, FP64) replacement for the explicit DG right-hand-side path of
pecos/tps (RHSoperator::Mult and callees).  The product is the C-ABI shared library
``tps_b200/lib/libtpsb200.so`` (``include/tpsb200.h``); this package is only the ctypes binding that
tests, ``bench.py`` and Python drivers (the reference ships ``src/tps.py`` style drivers too) use.
There is no CPU fallback: importing works anywhere, every compute call needs a CUDA device."""
from .capi import (TpsbError, BcDesc, LteTables, Physics, PlasmaModels, RhsOperator, build_library, cartesian_hex_mesh, cartesian_hex_partition, cartesian_quad_mesh, host_pipe_schedule,
                   cylinder_ogrid_mesh, quad_box_face_attr, partition_elements, partition_mesh,
                   lib, library_path, make_halo_desc)  # noqa: F401

__all__ = ["TpsbError", "BcDesc", "LteTables", "Physics", "PlasmaModels", "RhsOperator", "build_library", "cartesian_hex_mesh", "cartesian_hex_partition", "cartesian_quad_mesh", "host_pipe_schedule",
           "cylinder_ogrid_mesh", "quad_box_face_attr", "partition_elements", "partition_mesh",
           "make_halo_desc", "lib", "library_path"]
