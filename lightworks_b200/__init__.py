"""
lightworks_b200 -- B200 (sm_100a) implementation of the Lightworks emulator probability hot path.

Public surface (mirrors lightworks.emulator.backends):
    PermanentBackend, SLOSBackend, Backend   drop-in backend classes (backends.py)
    install()                                 redirect lightworks.emulator.Backend to them
    Context, default_context                  direct access to the C ABI (include/lwb200.h)
"""
from ._lib import (  # noqa: F401
    LWB_C64,
    LWB_C128,
    BackendError,
    Comm,
    Context,
    PhotonNumberError,
    DeviceDistribution,
    default_context,
    fock_count,
    fock_rank,
    fock_unrank,
    set_option,
    set_tuning,
    shard_cuts,
    slos_rank,
    slos_unrank,
    source_statistics,
)

from .backends import (  # noqa: E402, F401
    Backend,
    LazyDistribution,
    PermanentBackend,
    SLOSBackend,
    install,
    set_array_sampler,
    set_convolve_on_gpu,
    set_n_gpus,
    set_native_source_statistics,
    uninstall,
)
from .convolution import DistributionAlgebra, pdist_calc, sample_source_outputs  # noqa: E402, F401
from .state import CompiledCircuitView, State  # noqa: E402, F401

__version__ = "0.1.0"
