"""vln-imagine_b200: B200-native (sm_100a) implementation of the VLN-Imagine navigation hot path.

Python host modules mirror the reference's nn.Module API (DUET ``VLNBert`` / HAMT ``VLNBertCMT``);
every arithmetic step runs in hand-written CUDA kernels reached through the C-ABI library
``libvlnimagine.so`` (include/vlnimagine.h).  There is no CPU or eager-PyTorch fallback: importing
the compute modules without the built library raises.
"""
__version__ = '0.1.0'
