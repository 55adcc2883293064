from . import hmc, nuts  # noqa: F401
