"""Run the reference's OWN test files, unmodified, against bopy_b200 (authoring container only: needs /root/reference).

`bopy` and its sub-modules are aliased to `bopy_b200` in sys.modules, the absent third-party modules the reference's
tests import at module level get stand-ins (GPy: tests/gpy_standin.py -- the interface only; scipydirect: scipy's DIRECT
behind the same `minimize` signature; dppy / pyDOE / sobol_seq are not needed because bopy_b200 does not import them),
and pytest collects /root/reference/tests as they lie.

    python tools/run_reference_tests.py [-k expr]        # prints pytest's summary and a JSON line of the outcome

Without a CUDA device every test that fits a surrogate stops at `NativeLibraryError: no CUDA device is visible` (there
is no CPU fallback by design); the validation / shape / exception tests run to completion.  On the GPU box the tree has no
/root/reference: tests/test_gpu_api.py and tests/test_api_cpu.py restate the same cases (see REFERENCE_TEST_MAP there).
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/tests"


def alias_bopy():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gpy_standin
    gpy_standin.install()
    if "scipydirect" not in sys.modules:           # the reference's optimizer module is not imported: only tests might
        sd = types.ModuleType("scipydirect")
        sys.modules["scipydirect"] = sd
    import bopy_b200
    sys.modules["bopy"] = bopy_b200
    for name in ("acquisition", "bayes_opt", "benchmark_functions", "bounds", "callback", "exceptions", "initial_design",
                 "mixin", "optimizer", "surrogate"):
        sys.modules[f"bopy.{name}"] = getattr(bopy_b200, name)
    return bopy_b200


class Outcomes:
    def __init__(self):
        self.passed, self.failed, self.errors, self.no_device = [], [], [], []

    def pytest_runtest_logreport(self, report):
        if report.when == "call" and report.passed:
            self.passed.append(report.nodeid)
        elif report.failed:
            text = str(report.longrepr)
            bucket = self.no_device if ("no CUDA device is visible" in text or "NativeLibraryError" in text) else (
                self.failed if report.when == "call" else self.errors)
            bucket.append(report.nodeid)


def main(argv=None):
    import pytest
    if not os.path.isdir(REF_TESTS):
        print(json.dumps({"skipped": "no /root/reference on this machine"}))
        return 0
    alias_bopy()
    out = Outcomes()
    args = [REF_TESTS, "-q", "-p", "no:cacheprovider", "--rootdir", "/tmp", "-W", "ignore"] + list(argv or sys.argv[1:])
    pytest.main(args, plugins=[out])
    summary = {"passed": len(out.passed), "stopped_at_no_cuda_device": len(set(out.no_device)), "failed": sorted(set(out.failed)),
               "errors": sorted(set(out.errors))}
    print(json.dumps(summary))
    return summary


if __name__ == "__main__":
    main()
