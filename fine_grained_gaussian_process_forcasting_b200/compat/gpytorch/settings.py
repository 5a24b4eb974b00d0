from fine_grained_gaussian_process_forcasting_b200.gpcompat import (  # noqa: F401
    num_likelihood_samples, check_cholesky, variational_cholesky_jitter, min_variance)
