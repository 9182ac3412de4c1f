"""gcbfplus_b200 -- B200-native (sm_100a CUDA) hot path of GCBF+ behind the reference's
``gcbfplus.env`` / ``gcbfplus.algo`` / ``gcbfplus.trainer`` surface.

Host code is Python + PyTorch tensors (device memory, streams, torch.distributed);
all arithmetic of the hot path runs in libgcbf_b200.so (include/gcbf_b200.h).
"""
__version__ = "0.1.0"
