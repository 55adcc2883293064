from . import dynamic_hmc, hmc, integrators, nuts  # noqa: F401
