from fine_grained_gaussian_process_forcasting_b200.gpcompat import (DeepApproximateMLL, ExactMarginalLogLikelihood,  # noqa: F401
                                                                   VariationalELBO)
