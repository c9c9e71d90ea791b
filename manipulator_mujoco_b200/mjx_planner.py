"""Module-name shim: ``from mjx_planner import cem_planner`` (reference mpc_planner.py:2) keeps working
when this directory is on ``sys.path``."""
from .planner import cem_planner  # noqa: F401
