from .sequential import SequentialGraphNetwork, Module
from .embedding import (OneHotEncoding, RadialBasisEncoding, SphericalEncoding, Broadcast, RelativePositionEncoding,
                        symmetricCutoff)
from .pointwise import PointwiseLinear, TensorProductExpansion, Concat, LayerNormalization
from .message_passing import MessagePassing, FactorizedConvolution
from .scaling import PerTypeScaleShift
from .output import Pooling, GradientOutput
