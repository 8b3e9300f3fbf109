"""genhancer_b200 -- B200-native (sm_100a) implementation of GenHancer's stage-1/stage-2 training step.

Host side: PyTorch (device memory, streams, torch.distributed).  Compute: hand-written CUDA
kernels behind the C ABI in ``include/genhancer_b200.h`` (``libgenhancer_b200.so``).
"""
__version__ = "0.1.0"
