class VectorQuantizer2:
    pass
