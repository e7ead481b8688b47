"""Stand-in for python-fcl, used ONLY to import the unmodified reference in the build container.

The learned-sampler path constructs FCL objects (static_objects.py:31-36, obstacle.py:50-51,
fcl_utils.py:43-55) but never queries them; every collision on the path goes through
continuousCollideSpheres.  The baseline samplers do query ONE thing:
StaticObjects.collide_sphere -> DynamicAABBTreeCollisionManager.collide(sphere) against sphere obstacles
(static_objects.py:77-106).  FCL itself (0.5.0, Dockerfile:63) is absent, so that single query is restated here from its
published narrow phase `sphereSphereIntersect`: no collision iff ||c2 - c1|| > r1 + r2 (centres padded to 3-D, fp64).
Everything else stays inert.  Parity at this boundary is therefore UNPINNED.  Test infrastructure, not product code.
"""
import math


class _Inert:
    def __init__(self, *a, **k):
        self.args = a

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


class CollisionResult:
    def __init__(self, *a, **k):
        self.is_collision = False


class CollisionData:
    def __init__(self, request=None, *a, **k):
        self.request = request
        self.result = CollisionResult()


class ContinuousCollisionRequest(_Inert):
    pass


class ContinuousCollisionResult(_Inert):
    pass


class DynamicAABBTreeCollisionManager:
    def __init__(self, *a, **k):
        pass

    def registerObjects(self, objs):
        self.objs = list(objs)

    def collide(self, obj, rdata, callback=None):
        """sphere vs the registered spheres (CollisionObject(Sphere(rad), Transform(pos3)))"""
        r1, c1 = obj.args[0].args[0], obj.args[1].args[0]
        for o in getattr(self, "objs", []):
            r2, c2 = o.args[0].args[0], o.args[1].args[0]
            d = [float(c2[k]) - float(c1[k]) for k in range(3)]
            if not (math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > r1 + r2):
                rdata.result.is_collision = True
                return

    def setup(self):
        pass


def defaultCollisionCallback(*a, **k):
    raise RuntimeError("fcl stub")


def collide(*a, **k):
    raise RuntimeError("fcl stub")


def continuousCollide(*a, **k):
    raise RuntimeError("fcl stub")
