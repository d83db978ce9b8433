from micmec_b200.chk import load_chk, dump_chk  # noqa: F401
