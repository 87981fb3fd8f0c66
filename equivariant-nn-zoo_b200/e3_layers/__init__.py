"""Drop-in mirror of the reference's ``e3_layers`` package for the accelerated hot path:
same module / config API (irreps strings, data keys, ``config_*`` registry), B200 kernels
underneath (``e3b200`` + ``libe3b200.so``)."""
