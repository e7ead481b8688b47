"""ctrm_b200 — B200-native construction of CTRM timed roadmaps (the hot path of omron-sinicx/ctrm).

Public surface (mirrors the reference's names for this path):
  ctrm_b200.roadmap_learned.get_timed_roadmaps_multiple_paths_with_learned_indicator
  ctrm_b200.environment.Instance / ObstacleSphere / StaticObjects / continuous_collide_spheres
  ctrm_b200.timed_roadmap.TimedNode / TimedRoadmap / CSRRoadmaps
  ctrm_b200.planner.PrioritizedPlanning / Result   (the roadmaps' consumer)
  ctrm_b200.dataset.demonstration_features          (training-time input features)
  ctrm_b200.shims.{sphere_collision_check_wrapper, cost_to_go_wrapper}   pybind11-compatible modules
Everything computes on the GPU through libctrm_b200.so (include/ctrm_b200.h); importing the package does not need
a GPU, calling it does.
"""
from .environment import Instance, ObstacleSphere, StaticObjects, continuous_collide_spheres, load_instance  # noqa: F401
from .roadmap_learned import (PipelinedBuilder, RoadmapBuilder, get_builder,  # noqa: F401
                              get_timed_roadmaps_multiple_paths_with_learned_indicator)
from .planner import PrioritizedPlanning, Result  # noqa: F401
from .timed_roadmap import CSRRoadmaps, LazyTimedRoadmap, TimedNode, TimedRoadmap  # noqa: F401

__all__ = ["Instance", "ObstacleSphere", "StaticObjects", "continuous_collide_spheres", "load_instance", "RoadmapBuilder", "PipelinedBuilder",
           "get_builder", "get_timed_roadmaps_multiple_paths_with_learned_indicator", "CSRRoadmaps", "LazyTimedRoadmap",
           "TimedNode", "TimedRoadmap", "PrioritizedPlanning", "Result"]
