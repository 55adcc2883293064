from . import dynamic_hmc, ghmc, hmc, integrators, nuts  # noqa: F401
from . import metrics  # noqa: F401,E402
