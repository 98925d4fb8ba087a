"""morphik-core_b200 -- B200-native ColPali late-interaction (MaxSim) path for Morphik.

Host side is Python (like the reference) calling hand-written sm_100a CUDA through the C-ABI in
include/b200ms.h (ctypes).  Modules:
  _native   ctypes binding of libb200ms.so (fails loudly when the library is missing; there is no CPU fallback)
  index     MaxSimIndex: device-resident packed corpus + search (torch only holds the tensors)
  models    DocumentChunk, the record type of core/models/chunk.py
  store     B200MultiVectorStore, the BaseVectorStore plugin (core/vector_store/base_vector_store.py:7-65)
  sharded   document-sharded multi-GPU search (torch.distributed / NCCL all-gather of per-shard top-k)
"""
__version__ = "0.1.0"
