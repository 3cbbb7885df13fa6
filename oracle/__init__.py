"""CPU oracle for the voxelization hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (molvoxel_b200/) never does.
"""
from .oracle import OracleVoxelizer, build_oracle, oracle_forward_batch  # noqa: F401
