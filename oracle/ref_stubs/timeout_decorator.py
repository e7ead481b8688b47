"""Inert stand-in for timeout_decorator (planner only; off the hot path)."""


class TimeoutError(Exception):
    pass


def timeout(*a, **k):
    def deco(f):
        return f
    return deco
