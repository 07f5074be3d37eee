"""stellarscope_b200 -- B200-native EM read-reassignment path of Stellarscope.

Replaces, behind the reference's own Python model API, the hot path
``TelescopeLikelihood`` / ``_fit_pooling_model`` / ``reassign`` / count
aggregation (reference ``stellarscope/utils/model.py:87-507, 671-804,
1239-1345``) with hand-written sm_100a CUDA kernels reached through a C ABI
(``include/stellarscope_b200.h``).
"""
__version__ = '0.1.0'

import os as _os

# One EM fit launches ~20 independent kernels (lane-kernel size classes, cluster kernels) on as many
# streams.  The driver maps streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8);
# streams that share a queue serialise, which would run the fit in two or three waves.  The variable
# is read when the CUDA context is created, so it is set at import -- before torch touches the GPU
# -- unless the user chose a value.
_os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
