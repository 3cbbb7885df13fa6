"""molvoxel_b200 — B200-native voxelization backend behind the molvoxel API.

    import molvoxel_b200 as molvoxel
    vox = molvoxel.create_voxelizer(resolution=0.5, dimension=64, library="b200")
    grid = vox.forward_types(coords, center, types, radii=1.0)      # (C, 64, 64, 64) CUDA tensor

Factory signature and defaults follow reference molvoxel/__init__.py:9-40.  The reference hard-codes
its three library names (`assert library in [...]`, :33), so this package ships the factory with the
extra name "b200"; the other names are delegated to an installed molvoxel if there is one.
"""
from __future__ import annotations

from .transform import RandomTransform, T
from .voxelizer import Voxelizer
from .sharding import shard_bounds, shard_batch, gather_grids
from .pointcloud import PointCloud, Collator, collate, mol_point_cloud, system_point_cloud
from .dx import write_grid_to_dx_file
from .sparse import SparseGrids

__version__ = "0.2.0"
__all__ = ["create_voxelizer", "create_random_transform", "Voxelizer", "RandomTransform", "T",
           "shard_bounds", "shard_batch", "gather_grids", "PointCloud", "Collator", "collate", "mol_point_cloud",
           "system_point_cloud", "write_grid_to_dx_file", "SparseGrids"]


def create_random_transform(random_translation: float = 0.0, random_rotation: bool = False,
                            library: str = "b200", **kwargs) -> RandomTransform:
    if library == "b200":
        return RandomTransform(random_translation, random_rotation)
    import molvoxel  # delegate numpy / numba / torch to the reference package when it is installed
    return molvoxel.create_random_transform(random_translation, random_rotation, library, **kwargs)


def create_voxelizer(resolution: float = 0.5, dimension: int = 64, radii_type: str = "scalar",
                     density_type: str = "gaussian", library: str = "b200", **kwargs) -> Voxelizer:
    assert library in ["b200", "numba", "numpy", "torch"]
    if library == "b200":
        return Voxelizer(resolution, dimension, radii_type, density_type, **kwargs)
    import molvoxel
    return molvoxel.create_voxelizer(resolution, dimension, radii_type, density_type, library, **kwargs)
