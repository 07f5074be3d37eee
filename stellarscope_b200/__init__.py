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

# Host helper threads (csrc/ss_hostutil.c walks the reference's containers and cuts rows out of the matrix with a
# few threads; engine.py copies through a pinned ring with a few more).  With one process per GPU every rank runs
# them at the same time: 8 ranks x 16 threads on a 32-core host was the limiter of the round-1 scaling run.  Under
# torchrun the per-process default becomes cores / local world size (at least 2) unless the user chose a value.
try:
    _lw = int(_os.environ.get('LOCAL_WORLD_SIZE', '1'))
except ValueError:
    _lw = 1
if _lw > 1:
    _os.environ.setdefault('SS_HOST_THREADS', str(max(2, min(16, (_os.cpu_count() or 2) // _lw))))
