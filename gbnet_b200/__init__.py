"""gbnet_b200 — B200-native Gibbs sampler for gbnet's OR-NOR transcriptional-logic network.

``ModelORNOR`` keeps the reference's Python API (gbnet/ModelORNOR.pyx); ``Sampler`` is the
integer-coded handle over the C ABI (include/gbnet_b200.h).  The CUDA library is loaded on first
use; there is no CPU fallback.

Names are resolved on first access (PEP 562): importing the package - or only ``gbnet_b200.synth``, as
the reference arm of bench.py does - maps neither the CUDA library nor the Cython extension.
"""
import importlib
import sys
import types

__version__ = "0.2.0"

_LAZY = {
    # Cython binding of the C ABI (gbnet_b200/ModelORNOR.pyx, built by build.build_cython()); None when it is not built
    "PyModelORNOR": ("ModelORNOR", "PyModelORNOR"),
    "KERNEL_AUTO": ("engine", "KERNEL_AUTO"), "KERNEL_FAST": ("engine", "KERNEL_FAST"),
    "KERNEL_STRICT": ("engine", "KERNEL_STRICT"), "GbnError": ("engine", "GbnError"),
    "Sampler": ("engine", "Sampler"), "device_count": ("engine", "device_count"),
    "philox_kat": ("engine", "philox_kat"),
    "gen_network": ("synth", "gen_network"), "redraw_evidence": ("synth", "redraw_evidence"),
    "shape": ("synth", "shape"), "cli_default_hyper": ("synth", "cli_default_hyper"),
}


def __getattr__(name):
    try:
        module, attr = _LAZY[name]
    except KeyError:
        raise AttributeError(f"module 'gbnet_b200' has no attribute {name!r}") from None
    if name == "PyModelORNOR":
        try:
            ext = importlib.import_module(".ModelORNOR", __name__)
        except ImportError:
            value = None
        else:
            value = ext.PyModelORNOR
    else:
        value = getattr(importlib.import_module("." + module, __name__), attr)
    globals()[name] = value
    return value


def __dir__():
    return sorted(list(globals()) + list(_LAZY) + ["ModelORNOR"])


class _Package(types.ModuleType):
    """``gbnet_b200.ModelORNOR`` is the class (as ``gbnet.ModelORNOR`` is in the reference, gbnet/__init__.py:1),
    also after the import system has bound the extension module of the same name as a submodule."""

    @property
    def ModelORNOR(self):
        return importlib.import_module(".model", __name__).ModelORNOR

    @ModelORNOR.setter
    def ModelORNOR(self, value):
        pass


sys.modules[__name__].__class__ = _Package
