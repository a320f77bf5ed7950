"""B200-native vector-quantisation bottleneck for VQ-W-Net (drop-in for the quantiser of
Kaz-K/medical-image-editing, reference `src/networks/vq`).

Layout mirrors the reference's `src/` tree so the pieces drop into it unchanged:

    src/networks/vq/      VQ (= VQModule): same constructor / forward / lookup / get_codebook / buffers
    src/functions/        VQFunction: custom autograd Function calling the C-ABI (ctypes); EmbeddingLoss (cross-view cluster loss)
    src/trainers/         plain torch.distributed data-parallel trainer for VQ-W-Net
    src/utils/            get_world_size / is_distributed (the only utils the hot path uses)
    csrc/                 hand-written CUDA for sm_100a + the C-ABI (include/vq_b200.h)

There is no CPU fallback: every op raises if the CUDA library is missing or a tensor is not on a GPU.
"""
from ._native import lib, lib_path, build_native  # noqa: F401
from .src.networks.vq import VQ, VQModule  # noqa: F401
from .src.functions.vq_function import VQFunction, vq_lookup  # noqa: F401
from .src.functions.embed_loss import EmbeddingLoss  # noqa: F401

__all__ = ["VQ", "VQModule", "VQFunction", "vq_lookup", "EmbeddingLoss", "lib", "lib_path", "build_native"]
