from ctrm_b200.environment import *  # noqa: F401,F403
from ctrm_b200.environment import Instance, ObstacleSphere, StaticObjects, continuous_collide_spheres  # noqa: F401
