"""agent_bom_b200 — B200-native blast-radius engine for agent-bom's exposure-graph hot path.

Layers:
  csrc/ + include/abb200.h   hand-written sm_100a kernels behind a C ABI (libabb200.so)
  engine.DeviceGraph         ctypes face of that ABI (host-buffer calls)
  torch_api / dist           device-buffer calls on torch tensors; NCCL replication + source sharding
  graph.*                    the reference's Python graph surface (UnifiedGraph, compute_dependency_reach, derived paths)
  store / backend            GraphStoreProtocol / GraphBackend drop-ins
"""

__version__ = "0.1.0"
