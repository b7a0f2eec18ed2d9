"""fsnerf_b200 — B200-native (sm_100a) implementation of the fs-nerf ray-march
training/render hot path, a drop-in behind the reference's own entry points:

    fsnerf_b200.render.rendering   render_rays / render_frame / render_path
    fsnerf_b200.core.models        NeRF / PositionalEncoder
    fsnerf_b200.utils.utilities    get_rays / to_ndc / get_chunks

All arithmetic runs in hand-written CUDA kernels (libfsnerf_b200.so, C ABI in
include/fsnerf_b200.h); PyTorch is only device-memory / stream / NCCL plumbing.
There is no CPU fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
