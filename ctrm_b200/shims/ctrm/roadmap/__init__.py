from ctrm_b200.roadmap import (get_common_roadmaps, get_fixed_structured_roadmaps, get_grid, get_random_samples,  # noqa: F401
                               get_timed_roadmaps_fully_random, get_timed_roadmaps_grid, get_timed_roadmaps_grid_common,
                               get_timed_roadmaps_grid_common_2d_fast, get_timed_roadmaps_random,
                               get_timed_roadmaps_random_common, valid_move)
from ctrm_b200.timed_roadmap import TimedNode, TimedRoadmap  # noqa: F401
