"""oac_explore_b200: B200-native (sm_100a) implementation of the per-gradient-step hot path of
amarildolikmeta/oac-explore -- GPU-resident replay gather, fused OAC / P-OAC / G-OAC updates and
the optimistic exploration step -- behind the reference's own Python interfaces.

The package directory is named ``oac_explore_b200`` (a hyphen is not importable).
The compute lives in ``liboac_b200.so`` (C ABI: include/oac_b200.h); importing the trainers
without that library, or without a CUDA device, raises.
"""
from ._lib import lib, LIB_PATH  # noqa: F401

__all__ = ["lib", "LIB_PATH"]
