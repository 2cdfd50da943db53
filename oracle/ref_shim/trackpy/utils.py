def validate_tuple(value, ndim):
    """Scalar -> ndim-tuple; sequence of length ndim -> tuple; else ValueError."""
    if not hasattr(value, '__iter__'):
        return (value,) * ndim
    if len(value) == ndim:
        return tuple(value)
    raise ValueError("List length should have same length as image dimensions.")
