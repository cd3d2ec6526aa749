"""Annotation-only stand-in for the `torchtyping` package the reference imports
but this image lacks (setup.py pins torchtyping==0.1.4).  Used ONLY by
tests/golden/make_golden.py to import the unmodified reference in-container."""


class TensorType:
    def __class_getitem__(cls, item):
        return cls
