def load_obj(*a, **k):
    raise NotImplementedError
