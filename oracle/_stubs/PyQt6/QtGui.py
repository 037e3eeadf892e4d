def __getattr__(name):
    return type(name, (), {})
