from fine_grained_gaussian_process_forcasting_b200.gpcompat import DeepGPLayer, DeepGP  # noqa: F401
