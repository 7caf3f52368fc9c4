"""bitdelta_b200 -- the BitDelta W1A16 hot path as hand-written sm_100a CUDA behind a C ABI.

Python surface mirrors the reference (FasterDecoding/BitDelta):
    bitdelta_b200.binary_gemm_kernel : pack, unpack, binary_matmul, binary_bmm
    bitdelta_b200.diff               : BinaryDiff, compress_diff, save_diff, load_diff, save_full_model
    bitdelta_b200.demo_backend       : DiffCompressModule, DataParallelModule, register/unregister_diff_compress, DiffCompress
    bitdelta_b200.decode             : greedy_steps, greedy_decode, streaming_generator (demo_backend.py:190-258 without the server)
"""
from . import _lib  # noqa: F401  (fails loudly if libbitdelta_b200.so has not been built)
from .binary_gemm_kernel import binary_bmm, binary_matmul, pack, unpack
from .demo_backend import (
    DataParallelModule,
    DiffCompress,
    DiffCompressModule,
    fuse_sibling_projections,
    group_projections,
    load_checkpoints_stacked,
    register_diff_compress,
    unregister_diff_compress,
)
from .diff import BinaryDiff, compress_diff, fold_into, load_diff, save_diff, save_full_model
from . import parallel  # noqa: F401
from .decode import GraphedDecoder, greedy_decode, greedy_steps, streaming_generator

__all__ = [
    "pack", "unpack", "binary_matmul", "binary_bmm",
    "BinaryDiff", "compress_diff", "save_diff", "load_diff", "save_full_model", "fold_into",
    "DiffCompressModule", "DataParallelModule", "register_diff_compress", "unregister_diff_compress", "DiffCompress", "group_projections", "fuse_sibling_projections", "load_checkpoints_stacked",
    "greedy_steps", "greedy_decode", "streaming_generator", "GraphedDecoder",
]
