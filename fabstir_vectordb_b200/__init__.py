"""fabstir_vectordb_b200 — B200-native engine for fabstir-vectordb's search hot path.

Layout:
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/fvdb.h) -> libfvdb_b200.so
  _lib.py    ctypes loader (fails loudly when the .so is missing; no CPU fallback)
  engine.py  numpy-facing wrapper over the C ABI
  index.py   host mirror of the reference API: IVFIndex / HNSWIndex / HybridIndex
  chunk.py   VectorChunk CBOR <-> dense arrays (include/fvdb_chunk.h)
  synth.py   numpy twin of the device data generator
  shard.py   multi-GPU list-sharded search driver (torch.distributed / NCCL plumbing)
"""
from . import _lib  # noqa: F401
from .engine import (DimensionMismatch, DuplicateVector, Engine, FvdbError,  # noqa: F401
                     InconsistentDimensions, InsufficientTrainingData, InvalidConfig, NanInput,
                     NoDevice, NotTrained, PinnedArray, VectorNotFound)
from .chunk import ChunkError, VectorChunk, decode_vector_chunk, encode_vector_chunk  # noqa: F401
from .index import (AddClustersResult, BalanceResult, ClusterStats, HNSWConfig, HNSWIndex,  # noqa: F401
                    HybridConfig, HybridIndex, HybridSearchConfig, IVFConfig, IVFIndex, InvalidParameter,
                    NotInitialized, OptimizationResult, RetrainResult, SearchConfig,
                    SearchResult, TrainResult)

__all__ = [n for n in dir() if not n.startswith("_")]
