"""tfep_b200: B200 (sm_100a) implementation of the tfep MAF / (T)FEP hot path.

Drop-in for ``tfep.nn.flows.MAF`` (MADE conditioner + Affine / NeuralSpline / SOS / Moebius / Mixed
transformers, forward / inverse with log|det J|), ``tfep.analysis.fep_estimator`` and
``tfep.analysis.bootstrap``, behind the reference's own Python API.  All arithmetic runs in hand-written
CUDA kernels loaded through the C ABI of include/tfep_b200.h; there is no CPU fallback.
"""

__version__ = '0.1.0'
