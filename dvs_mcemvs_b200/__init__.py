"""B200-native DSI ray-voting engine: drop-in for the mapping hot path of tub-rip/dvs_mcemvs.

Layout:
  csrc/   hand-written sm_100a CUDA kernels + the C-ABI (include/emvs_b200.h) -> lib/libemvs_b200.so
  host/   C++ host classes with the reference's names (Grid3D, EMVS::MapperEMVS, ...) over the C-ABI
  api.py  the same mirror in Python (ctypes), used by tests/ and bench.py
  synth.py  seeded synthetic rigs / event streams for BASELINE.json's configs
"""
__all__ = ["api", "synth"]
