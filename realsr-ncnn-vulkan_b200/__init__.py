"""realsr-ncnn-vulkan_b200 -- B200-native drop-in for the RealSR 4x tile path of nihui/realsr-ncnn-vulkan.

The product is the C-ABI shared library built from csrc/ (include/realsr_b200.h); this package is the Python mirror
of the reference's `class RealSR` (reference: src/realsr.h:13-42) on top of it.  The directory name contains a '-',
so it is imported through the `realsr_ncnn_vulkan_b200` symlink at the repo root.
"""
from .realsr import RealSR, RealSRError, lib_path, load_library  # noqa: F401
from .sharded import process_sharded, split_tiles, tile_costs  # noqa: F401
