"""bullet3_b200 -- B200-native GPU rigid-body step behind Bullet3's b3GpuRigidBodyPipeline API.

The product is `libb3b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/b3b200.h) plus the C++ drop-in classes in csrc/host/.  The Python
modules here only bind that ABI for the tests and bench.py; nothing in this
package computes physics on the CPU.
"""
from . import capi  # noqa: F401
from .capi import World, Broadphase, B3Error, lib  # noqa: F401
