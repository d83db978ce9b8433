class _Periodic(dict):
    def __missing__(self, key):
        raise KeyError(key)
periodic = _Periodic()
