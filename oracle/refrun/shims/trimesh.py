"""Shim: kinematics-only URDFs never build meshes, the names just need to exist."""


class Trimesh:
    pass


class Scene:
    pass


class _NS:
    def __getattr__(self, k):
        raise NotImplementedError("trimesh shim: %s" % k)


creation = _NS()
transformations = _NS()
visual = _NS()
exchange = _NS()
util = _NS()


def load(*a, **k):
    raise NotImplementedError("trimesh shim: load")
