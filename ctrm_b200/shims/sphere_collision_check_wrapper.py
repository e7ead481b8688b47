"""GPU replacement of cpp_helper/sphere_collision_check_wrapper (collision_check.cpp:16-67): same module name,
function name, argument order and return type."""
from ctrm_b200.environment import continuous_collide_spheres as _impl


def continuousCollideSpheres(from1, to1, rad1, from2, to2, rad2) -> bool:
    """check collision of two moving spheres"""
    return _impl(from1, to1, rad1, from2, to2, rad2)
