from ctrm_b200.planner import PrioritizedPlanning, Result  # noqa: F401
