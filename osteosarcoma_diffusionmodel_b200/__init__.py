"""B200-native conditional-DDPM hot path of rare-resilience-ai/Osteosarcoma_DiffusionModel.

Host code is PyTorch (device memory, streams, torch.distributed); every kernel on the path is
hand-written CUDA for sm_100a behind the C-ABI in include/osteo_ddpm.h.
"""
__version__ = "0.1.0"
