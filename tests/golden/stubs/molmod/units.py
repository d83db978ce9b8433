from micmec_b200.units import *  # noqa: F401,F403
