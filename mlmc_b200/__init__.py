"""mlmc_b200 -- B200-native implementation of the GeoMop/MLMC estimation hot path.

Drop-in for ``mlmc.moments``, ``mlmc.quantity.quantity_estimate``, ``mlmc.estimator.Estimate`` and
``mlmc.tool.simple_distribution.SimpleDistribution`` (reference v1.0.2): same names, signatures and result
objects; the arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of ``include/mlmcb200.h``
(``mlmc_b200/_lib/libmlmcb200.so``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
