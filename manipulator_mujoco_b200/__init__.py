"""B200-native CEM planner for the UR5e + Hand-E scene (drop-in for the reference's
``sampling_based_planner/mjx_planner.py``).  ``cem_planner`` needs a CUDA device and the in-tree
``libcemk.so`` (built by ``__graft_entry__.build()``); importing the package itself does not."""

__all__ = ["cem_planner"]


def __getattr__(name):
    if name == "cem_planner":
        from .planner import cem_planner
        return cem_planner
    raise AttributeError(name)
