"""nextpolish_b200 — B200-native polishing engine behind NextPolish's lib/nextpolish1.so ABI.

csrc/     CUDA kernels (sm_100a), host I/O (BGZF/BAM/FASTA), the C ABI and the native CLI
lib/      built artefacts: nextpolish1.so (drop-in shared object), nextpolish1 (CLI)
binding   ctypes prototypes of include/nextpolish_b200.h
engine    Shard / Engine handle classes
nextpolish1  host-side mirror of the reference's lib/nextpolish1.py worker CLI
"""
__all__ = ["binding", "engine"]
