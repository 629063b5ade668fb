"""Exception classes of the path (reference: utils/errors.py)."""


class NanError(RuntimeError):
    pass


class NotPSDError(RuntimeError):
    pass


class CachingError(RuntimeError):
    pass
