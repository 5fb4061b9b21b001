"""localization_b200 — B200-native batched sliding-window LM solver for UWB range graphs.

Host-side mirror of the hot path of sair-lab/localization (`Localization::solve()`,
reference src/localization/localization.cpp:164-192).  Python here is tooling: it packs
windows, calls the C ABI of `libuwbgo.so` (CUDA, sm_100a) through ctypes and reads results
back.  Nothing in this package computes a solve on the CPU.
"""
from .graph import Topology, Batch, Config, Result  # noqa: F401
from .solver import Solver, UwbgoError  # noqa: F401

__all__ = ["Topology", "Batch", "Config", "Result", "Solver", "UwbgoError"]
