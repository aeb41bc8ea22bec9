"""B200-native DEFLATE compressor: drop-in for the compress path of
mparker97/deflate_compression.

The product is libdeflate_b200.so (hand-written sm_100a kernels + C host code);
this package is the Python mirror of its C interface."""
from .api import (  # noqa: F401
    Compressor,
    DflParams,
    DflError,
    build,
    compress_multi,
    deflate_compress,
    deflate_decompress,
    lib,
    lib_path,
)
from . import synth  # noqa: F401
