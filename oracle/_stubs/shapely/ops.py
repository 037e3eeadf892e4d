def nearest_points(*a, **k):
    raise NotImplementedError


def __getattr__(name):
    return type(name, (), {})
