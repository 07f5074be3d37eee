"""stellarscope_b200 -- B200-native EM read-reassignment path of Stellarscope.

Replaces, behind the reference's own Python model API, the hot path
``TelescopeLikelihood`` / ``_fit_pooling_model`` / ``reassign`` / count
aggregation (reference ``stellarscope/utils/model.py:87-507, 671-804,
1239-1345``) with hand-written sm_100a CUDA kernels reached through a C ABI
(``include/stellarscope_b200.h``).
"""
__version__ = '0.1.0'
