"""Inert stand-in for the OMPL/FCL SPARS helper (baseline sampler; out of scope)."""


def getSparsRoadmap2d(*a, **k):
    raise RuntimeError("spars_wrapper stub: out of scope")
