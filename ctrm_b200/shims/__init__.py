"""Drop-in modules with the names and call signatures of the reference's pybind11 helpers.  Put this directory
on sys.path (or `from ctrm_b200.shims import cost_to_go_wrapper`) in place of the compiled C++ modules."""
