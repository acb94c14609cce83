"""The Bayesian-optimisation loop (reference: bopy/bayes_opt.py:80-270) -- pure orchestration.

Control flow, event order and result records are the reference's; all numerical work happens in the
surrogate / acquisition / optimizer objects it is given."""
from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Tuple

import numpy as np

from .acquisition import AcquisitionFunction
from .bounds import Bounds
from .callback import Callback
from .initial_design import InitialDesign
from .optimizer import OptimizationResult, Optimizer
from .surrogate import Surrogate


@dataclass
class BOInitialDesignResult:
    """x_selected (n_init, d), f_selected (n_init,), incumbent x_opt_so_far (1, d) / f_opt_so_far."""

    x_selected: np.ndarray
    f_selected: np.ndarray
    x_opt_so_far: np.ndarray
    f_opt_so_far: float


@dataclass
class BOTrialResult:
    """x_selected (n_batch, d), f_selected (n_batch,), incumbent x_opt_so_far (1, d) / f_opt_so_far."""

    x_selected: np.ndarray
    f_selected: np.ndarray
    x_opt_so_far: np.ndarray
    f_opt_so_far: float


@dataclass
class BOResult:
    """Final incumbent plus the per-stage records."""

    x_opt: np.ndarray
    f_opt: float
    initial_design_result: BOInitialDesignResult
    trial_results: List[BOTrialResult]


class BayesOpt:
    """Sequential model-based minimisation of `objective_function` over `bounds`."""

    def __init__(self, objective_function: Callable[[np.ndarray], np.ndarray], surrogate: Surrogate,
                 acquisition_function: AcquisitionFunction, optimizer: Optimizer, initial_design: InitialDesign,
                 bounds: Bounds, callbacks: Optional[List[Callback]] = None):
        self.objective_function = objective_function
        self.surrogate = surrogate
        self.acquisition_function = acquisition_function
        self.optimizer = optimizer
        self.initial_design = initial_design
        self.bounds = bounds
        self.callbacks = callbacks
        self.x = np.array([])
        self.y = np.array([])

    # -- driver -------------------------------------------------------------------------------------
    def run(self, n_trials: int = 10, n_initial_design: int = 5) -> BOResult:
        design = self.run_initial_design(n_initial_design)
        trials = self.run_trials(n_trials)
        self.dispatch("on_bo_end", self)
        x_opt, f_opt = self.get_opt_so_far()
        return BOResult(x_opt=x_opt, f_opt=f_opt, initial_design_result=design, trial_results=trials)

    def run_initial_design(self, n_initial_design: int = 5) -> BOInitialDesignResult:
        x = self.initial_design.generate(self.bounds, n_initial_design)
        y = self.objective_function(x)
        self.append_to_dataset(x, y)
        self.dispatch("on_initial_design_end", self)
        self.surrogate.fit(self.x, self.y)
        self.acquisition_function.fit(self.x, self.y)
        x_best, f_best = self.get_opt_so_far()
        return BOInitialDesignResult(x_selected=x, f_selected=y, x_opt_so_far=x_best, f_opt_so_far=f_best)

    def run_trials(self, n_trials: int = 10) -> List[BOTrialResult]:
        return [self.run_trial() for _ in range(n_trials)]

    def run_trial(self) -> BOTrialResult:
        proposal = self.optimize_acquisition()
        x = proposal.x_min
        y = self.objective_function(x)
        self.append_to_dataset(x, y)
        self.update_surrogate()
        self.update_acquisition()
        self.dispatch("on_trial_end", self)
        x_best, f_best = self.get_opt_so_far()
        return BOTrialResult(x_selected=x, f_selected=y, x_opt_so_far=x_best, f_opt_so_far=f_best)

    # -- steps --------------------------------------------------------------------------------------
    def optimize_acquisition(self) -> OptimizationResult:
        result = self.optimizer.optimize()
        self.dispatch("on_acquisition_optimized", self, result)
        return result

    def update_surrogate(self) -> None:
        self.surrogate.fit(self.x, self.y)
        self.dispatch("on_surrogate_updated", self)

    def update_acquisition(self) -> None:
        self.acquisition_function.fit(self.x, self.y)
        self.dispatch("on_acquisition_updated", self)

    # -- plumbing -----------------------------------------------------------------------------------
    def dispatch(self, event: str, *args: Any) -> None:
        for callback in self.callbacks or ():
            getattr(callback, event)(*args)

    def append_to_dataset(self, x: np.ndarray, y: np.ndarray) -> None:
        if len(self.x) == 0 and len(self.y) == 0:
            self.x, self.y = x, y
        else:
            self.x = np.concatenate((self.x, x))
            self.y = np.concatenate((self.y, y))

    def get_opt_so_far(self) -> Tuple[np.ndarray, float]:
        best = int(np.argmin(self.y))
        return np.atleast_2d(self.x[best]), self.y[best]
