"""bgflow_b200 — a Blackwell (sm_100a) engine for bgflow's coupling-flow hot path.

The public names mirror ``bgflow``'s for the path this package accelerates: ``SequentialFlow``,
``SplitFlow`` / ``MergeFlow`` / ``SwapFlow`` / ``InverseFlow``, ``CouplingFlow``,
``AffineTransformer``, ``ConditionalSplineTransformer``, ``DenseNet``, ``WrapPeriodic``,
``GlobalInternalCoordinateTransformation``, ``BoltzmannGenerator``.  The arithmetic runs in
``libbgflow_b200.so`` (hand-written CUDA behind the C ABI of ``include/bgflow_b200.h``); the
package has no CPU fallback and raises if the library cannot be loaded.
"""

from .flows import (Flow, SequentialFlow, InverseFlow, SplitFlow, MergeFlow, SwapFlow, CouplingFlow,
                    WrapFlow, SetConstantFlow)
from .nets import DenseNet, MeanFreeDenseNet, WrapPeriodic
from .transformers import Transformer, AffineTransformer, ConditionalSplineTransformer
from .ic import (GlobalInternalCoordinateTransformation, RelativeInternalCoordinateTransformation,
                 MixedCoordinateTransformation, WhitenFlow)
from .cdf import (CDFTransform, DistributionTransferFlow, ConstrainGaussianFlow, TruncatedNormalDistribution,
                  SloppyUniform, MultiCDFFlow, MappedICTail, fuse_domain_maps)
from .bg import (BoltzmannGenerator, NormalDistribution, UniformDistribution, unnormalized_kl_div,
                 unormalized_nll, log_weights, log_weights_given_latent, effective_sample_size,
                 sampling_efficiency)
from .convert import from_reference
from . import engine, _lib, distributed

__version__ = "0.1.0"
