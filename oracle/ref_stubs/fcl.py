"""Inert stand-in for python-fcl, used ONLY to import the unmodified reference in the build container.

The learned-sampler path constructs FCL objects (static_objects.py:31-36, obstacle.py:50-51,
fcl_utils.py:43-55) but never queries them; every collision on the path goes through
continuousCollideSpheres. Test infrastructure, not product code.
"""


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        def _f(*a, **k):
            raise RuntimeError(f"fcl stub: {name} is not on the learned-sampler path")
        return _f


class Sphere(_Inert):
    pass


class Box(_Inert):
    pass


class Capsule(_Inert):
    pass


class Transform(_Inert):
    pass


class CollisionObject(_Inert):
    pass


class CollisionRequest(_Inert):
    pass


class CollisionResult(_Inert):
    pass


class CollisionData(_Inert):
    pass


class ContinuousCollisionRequest(_Inert):
    pass


class ContinuousCollisionResult(_Inert):
    pass


class DynamicAABBTreeCollisionManager:
    def __init__(self, *a, **k):
        pass

    def registerObjects(self, objs):
        pass

    def setup(self):
        pass


def defaultCollisionCallback(*a, **k):
    raise RuntimeError("fcl stub")


def collide(*a, **k):
    raise RuntimeError("fcl stub")


def continuousCollide(*a, **k):
    raise RuntimeError("fcl stub")
