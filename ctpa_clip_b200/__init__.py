"""ctpa_clip_b200 — B200-native (sm_100a) implementation of the CT-CLIP contrastive hot path.

Host code is Python/PyTorch (allocator, streams, torch.distributed); all arithmetic on the path runs in
hand-written CUDA behind the C-ABI of libctclip_sm100.so (include/ctclip_b200.h). No CPU fallback.
"""
__version__ = "0.1.0"
