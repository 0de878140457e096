"""gbnet_b200 — B200-native Gibbs sampler for gbnet's OR-NOR transcriptional-logic network.

``ModelORNOR`` keeps the reference's Python API (gbnet/ModelORNOR.pyx); ``Sampler`` is the
integer-coded handle over the C ABI (include/gbnet_b200.h).  The CUDA library is loaded on first
use; there is no CPU fallback.
"""
try:                                   # Cython binding of the same C ABI (gbnet_b200/ModelORNOR.pyx, built by
    from .ModelORNOR import PyModelORNOR   # build.build_cython()); the name below is rebound to the class, as the
except ImportError:                    # reference's __init__ does with its own extension module
    PyModelORNOR = None
from .engine import KERNEL_AUTO, KERNEL_FAST, KERNEL_STRICT, GbnError, Sampler, device_count, philox_kat  # noqa: F401
from .model import ModelORNOR  # noqa: F401
from .synth import gen_network, redraw_evidence, shape, cli_default_hyper  # noqa: F401

__version__ = "0.1.0"
