class Meshes:
    pass
