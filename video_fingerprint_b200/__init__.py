"""video_fingerprint_b200 - B200 (sm_100a) implementation of the video-fingerprint duplicate-detection hot path.

Public surface (mirrors /root/reference/model.py and /root/reference/fingerprint.py for the hot path only):

    create_model, VideoFingerprintAttention          reference model factory / module (state_dict compatible)
    VideoFingerprintScanner                          find_duplicates / save_results / per-video semantics
    threshold_join, topk_inner_product               the two similarity-search primitives
    sharding                                         multi-GPU partitioning + NCCL all-gather join
    compute_retrieval_metrics, compute_discrimination_metrics   the trainer's evaluation metrics over the same N x N scores

All device work goes through libvfp_b200.so (include/vfp_b200.h); there is no CPU or PyTorch fallback.
"""
from .model import VideoFingerprint3D, VideoFingerprintAttention, create_model  # noqa: F401
from .fingerprint import (  # noqa: F401
    VideoFingerprintScanner,
    duplicate_pairs,
    group_pairs_direct,
    group_pairs_topk,
    preprocess_frames_device,
    threshold_join,
    threshold_join_device,
    topk_inner_product,
    topk_inner_product_device,
)
from .metrics import compute_discrimination_metrics, compute_retrieval_metrics  # noqa: F401

__all__ = [
    "preprocess_frames_device",
    "compute_retrieval_metrics",
    "compute_discrimination_metrics",
    "create_model",
    "VideoFingerprintAttention",
    "VideoFingerprint3D",
    "VideoFingerprintScanner",
    "threshold_join",
    "threshold_join_device",
    "topk_inner_product",
    "topk_inner_product_device",
    "duplicate_pairs", "group_pairs_direct",
    "group_pairs_topk",
]
