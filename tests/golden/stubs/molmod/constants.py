from micmec_b200.units import boltzmann  # noqa: F401
