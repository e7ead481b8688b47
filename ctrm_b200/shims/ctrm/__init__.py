"""Importable `ctrm` alias for the paths ctrm_b200 replaces: put `ctrm_b200/shims` on sys.path (before or instead of the
reference's `src/`) and `ctrm.roadmap_learned`, `ctrm.environment` (incl. `.instance`, `.obstacle`, `.static_objects`, so
that the reference's `*_ins.pkl` pickles load with plain `pickle.load`), `ctrm.roadmap` and `ctrm.planner` resolve to the
GPU-backed implementations — hydra `_target_: ctrm.roadmap_learned.get_timed_roadmaps_multiple_paths_with_learned_indicator`
(scripts/conf/eval/*.yaml, scripts/eval.py:59) then needs no edit.  Nothing else of the reference's package is provided."""
