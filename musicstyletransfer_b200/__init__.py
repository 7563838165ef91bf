"""musicstyletransfer_b200 — B200-native (sm_100a) hot path of slyforce/MusicStyleTransfer.

Host side is Python (as the reference is), PyTorch is used for device memory, streams, autograd
ordering and torch.distributed only; every device computation is a hand-written CUDA kernel in
``csrc/`` reached through the C ABI of ``libmsx.so`` (``include/msx.h``).  There is no CPU fallback:
loading ``musicstyletransfer_b200.lib`` without a built library raises.
"""
__version__ = "0.1.0"
