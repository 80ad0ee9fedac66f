"""galaxysql_b200 — B200-native MPP operator hot path for GalaxySQL (hash join, hash aggregation, hash-partition
exchange) behind the C-ABI of include/gsql_gpu.h.  No CPU fallback: importing the compute API without the built
extension or without a CUDA device raises."""
__all__ = ["api", "native", "build", "synth"]
