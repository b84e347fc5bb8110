"""vndecorrelate_b200 — the velvet-noise decorrelation hot path of ckonst/VNDecorrelate on
B200 (sm_100a) CUDA kernels, behind the reference's Python API.

    from vndecorrelate_b200.decorrelation import VelvetNoise, HaasEffect, SignalChain
    from vndecorrelate_b200.optimization import optimize_velvet_noise

The native library (``_lib/libvnd_b200.so``, C ABI in ``include/vnd_b200.h``) is loaded on first
use; there is no CPU fallback.
"""

__version__ = "0.1.0"
