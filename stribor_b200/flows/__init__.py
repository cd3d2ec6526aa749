from .affine import *
from .coupling import *
from .spline import *
from .pointwise import *
