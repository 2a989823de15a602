"""cp360_b200 — B200 (sm_100a) spherical-projection hot path of CP-360-Weakly-Supervised-Saliency.

Host-side mirror of the reference's three operators, backed by libcp360.so (include/cp360.h):

    CubePad / CubePadding / get_pad_size     model/cube_pad.py
    Equi2Cube                                utils/equi_to_cube.py
    Cube2Equi                                utils/cube_to_equi.py

plus SphericalPipeline (pipeline.py), the batched, frame-sharded chain the benchmark measures.
"""
from . import _lib
from ._build import build_library, LIB_PATH
from .cube_pad import (autotune_cubepad, CubePad, CubePadding, get_pad_size, cubepad_forward, cubepad_index_map, cubepad_fused,
                       cubepad_cat, cubepad_bn_relu)
from .cam import SaliencyHead, cam_scores, cam_weight
from .cube_to_equi import Cube2Equi
from .equi_to_cube import Equi2Cube
from .hostmem import pinned_empty
from .io import backproject_files, load_cube_feat, load_npy, npy_header, save_npy
from .pipeline import (SphericalPipeline, TemporalCubePadSequence, gather_maps, gpu_numa_node, prefer_gpu_numa_node,
                       resnet50_cubepad_sites, shard_range)

__all__ = ["CubePad", "CubePadding", "get_pad_size", "cubepad_forward", "cubepad_index_map", "cubepad_fused", "cubepad_cat", "cubepad_bn_relu", "autotune_cubepad",
           "Equi2Cube", "Cube2Equi", "SaliencyHead", "cam_scores", "cam_weight", "SphericalPipeline", "TemporalCubePadSequence", "gather_maps", "resnet50_cubepad_sites",
           "shard_range", "build_library", "LIB_PATH", "load_npy", "save_npy", "npy_header", "load_cube_feat",
           "backproject_files", "pinned_empty"]
