from . import hmc, integrators, nuts  # noqa: F401
