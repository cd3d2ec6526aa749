"""stribor_b200 -- B200-native drop-in for the coupling-flow hot path of mbilos/stribor.

    import stribor_b200 as st
    flow = st.NormalizingFlow(st.UnitNormal(d), [st.Coupling(st.Spline(d, 16, latent_net=
               st.net.MLP(d, [64], d * 47), spline_type='quadratic'), mask='ordered_right_half'), ...])
    flow.cuda().log_prob(y)

Same class names, constructor arguments, mask strings and state-dict keys as the reference for
``NormalizingFlow / NeuralFlow / Coupling / ContinuousAffineCoupling / Affine / Spline /
net.MLP / net.TimeLinear / UnitNormal / util.get_mask``; the arithmetic runs in hand-written
sm_100a CUDA kernels behind a C ABI (``include/stribor_b200.h``).  CUDA tensors only.
"""
from .dist import *
from .flow import *
from .flows import *
from . import util
from . import net
from . import _ops
from .flows._native import invalidate_packed

__version__ = '0.1.0'
