"""B200-native (sm_100a) implementation of the GNGF training hot path of
FedeMont/collision_handling_in_instantNGP: `GeneralNeuralGaugeFields` forward + backward, in which the
`HashProbDistribution` MLP and `DifferentiableTopk` replace Instant-NGP's spatial hash.

    from collision_handling_in_instantngp_b200.models import GeneralNeuralGaugeFields

The classes keep the reference's signatures (see models.py); the arithmetic runs in libgngf_sm100.so
(include/gngf.h).  There is no CPU path: constructing a model without a CUDA device or without the built
library raises ``GngfError``.
"""
from ._lib import GngfError, LIB_PATH, launch_count, load  # noqa: F401

__version__ = "0.1.0"
