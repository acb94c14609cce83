"""Exception types of the drop-in API (reference: bopy/exceptions.py:1-2)."""


class NotFittedError(Exception):
    """Raised when predict / evaluate is called on an object that has not been fitted."""


class NativeLibraryError(RuntimeError):
    """The sm_100a extension is missing, failed to load, or returned an error code.

    There is deliberately no CPU fallback behind the B200 surrogate: when the CUDA path cannot
    run, this is raised."""
