"""Warning categories raised by the Krylov path (reference: utils/warnings.py)."""


class NumericalWarning(RuntimeWarning):
    """A numerical issue that may affect accuracy (e.g. CG stopped above its tolerance)."""


class PerformanceWarning(RuntimeWarning):
    """A slow code path was taken."""


class OldVersionWarning(UserWarning):
    pass
