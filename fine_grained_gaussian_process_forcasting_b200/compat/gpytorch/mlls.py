from fine_grained_gaussian_process_forcasting_b200.gpcompat import DeepApproximateMLL, VariationalELBO  # noqa: F401
