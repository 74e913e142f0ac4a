"""moleengine_b200 — a B200-native (sm_100a) physics tick for Starframe (MoleTrooper/moleengine).

Only the hot path `PhysicsWorld::tick` + cast/point queries is implemented, behind the C ABI in
include/starframe_gpu.h. `world.PhysicsWorld` mirrors the reference's Rust-facing API on top of it.
"""
from . import abi  # noqa: F401
from .world import GpuWorld, SfgError  # noqa: F401
