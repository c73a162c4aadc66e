"""mvoc_b200 — B200-native (sm_100a) implementation of the MVOC composition hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed) over the
C-ABI library ``mvoc_b200/lib/libmvoc_b200.so`` declared in ``include/mvoc_b200.h``.
"""
__version__ = "0.1.0"
