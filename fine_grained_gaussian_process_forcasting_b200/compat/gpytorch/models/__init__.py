from fine_grained_gaussian_process_forcasting_b200.gpcompat import GP, ApproximateGP, ExactGP  # noqa: F401
from . import deep_gps  # noqa: F401
