from fine_grained_gaussian_process_forcasting_b200.gpcompat import Mean, ConstantMean, LinearMean  # noqa: F401
