from ctrm_b200.timed_roadmap import TimedNode, TimedRoadmap  # noqa: F401
