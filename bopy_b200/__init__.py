"""bopy_b200 -- B200-native implementation of bopy's GP posterior -> acquisition -> argmin path.

Same public surface as tompretty/bopy (surrogate / acquisition / optimizer / bayes_opt / callback /
bounds / initial_design / benchmark_functions); the posterior and the acquisition functions run in
hand-written sm_100a kernels behind the C ABI in include/bopy_b200.h.
"""
from . import (acquisition, bayes_opt, benchmark_functions, bounds, callback, exceptions, initial_design, mixin,
               optimizer, surrogate)

__version__ = "0.1.0"
__all__ = ["acquisition", "bayes_opt", "benchmark_functions", "bounds", "callback", "exceptions",
           "initial_design", "mixin", "optimizer", "surrogate"]
