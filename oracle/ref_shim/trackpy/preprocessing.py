def bandpass(*args, **kwargs):
    raise NotImplementedError


def scalefactor_to_gamut(*args, **kwargs):
    raise NotImplementedError


def scale_to_gamut(*args, **kwargs):
    raise NotImplementedError
