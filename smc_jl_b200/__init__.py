"""smc_jl_b200 -- B200-native SMC particle engine behind SMC.jl's `smc(...)` / `Cloud` API.

The compute path is the CUDA library `libsmcb200.so` (C ABI in include/smcb200.h); this package
is the host-side mirror of the reference's Julia interface.  There is no CPU fallback.
"""
from .model import (AnSchorfheideLogLik, Beta, CAPMLogLik, Gamma, GaussRegLogLik, InverseGamma,  # noqa: F401
                    LinearEquationsLogLik, LinearGaussianLogLik, ModelSpec, Normal, Parameter, RootInverseGamma, Uniform,
                    make_spec, parameter)
from .cloud import Cloud  # noqa: F401,E402


def smc(*args, **kwargs):
    """smc(loglikelihood, parameters, data; ...) -- see smc_jl_b200.driver.smc (imports the CUDA library lazily)."""
    from .driver import smc as _smc
    return _smc(*args, **kwargs)


def get_cloud(path):
    """get_cloud(path) (src/util.jl:113-115): the `cloud` stored by `smc(...; savepath = path)`."""
    from .driver import load_cloud
    return load_cloud(path)[0]
