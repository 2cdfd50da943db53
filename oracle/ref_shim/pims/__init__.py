"""Minimal stand-in for `pims` (absent from this image).  TEST INFRASTRUCTURE ONLY."""
import numpy as np


class Frame(np.ndarray):
    """ndarray subclass that carries `frame_no` and `metadata`."""

    def __new__(cls, input_array, frame_no=None, metadata=None):
        obj = np.asarray(input_array).view(cls)
        obj.frame_no = frame_no
        obj.metadata = metadata if metadata is not None else {}
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.frame_no = getattr(obj, 'frame_no', None)
        self.metadata = getattr(obj, 'metadata', {})


class FramesSequence(object):
    def __getitem__(self, key):
        return self.get_frame(key)

    def __iter__(self):
        return (self.get_frame(i) for i in range(len(self)))


def pipeline(func):
    return func
