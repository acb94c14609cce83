"""Hooks fired by the BayesOpt loop (reference: bopy/callback.py:4-74).

Subclass `Callback` and override the events of interest; every hook receives the running
`BayesOpt` object (`on_acquisition_optimized` additionally the `OptimizationResult`)."""

EVENTS = (
    "on_initial_design_end",
    "on_acquisition_optimized",
    "on_surrogate_updated",
    "on_acquisition_updated",
    "on_trial_end",
    "on_bo_end",
)


class Callback:
    def on_initial_design_end(self, bo):
        """After the initial design has been evaluated (before the first fit)."""

    def on_acquisition_optimized(self, bo, opt_result):
        """After the optimizer proposed `opt_result` for the next trial."""

    def on_surrogate_updated(self, bo):
        """After the surrogate was refitted on the grown data set."""

    def on_acquisition_updated(self, bo):
        """After the acquisition function was refitted."""

    def on_trial_end(self, bo):
        """After one complete trial."""

    def on_bo_end(self, bo):
        """After the last trial."""
