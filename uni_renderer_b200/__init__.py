"""uni_renderer_b200 -- B200-native (sm_100a) implementation of the Uni-Renderer dual-stream denoising hot path.

Python host side mirroring the reference's module interface (models/controlnet.py, models/pipeline.py) on top of
hand-written CUDA kernels reached through the C ABI in include/unib200.h (libunib200.so, built in-tree by
uni_renderer_b200.build).  There is no CPU or PyTorch-eager fallback: importing the ops without the built library,
or running them without a CUDA device, raises.
"""
__version__ = "0.1.0"
