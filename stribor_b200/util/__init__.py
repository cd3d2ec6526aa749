from .mask import *
from .splines import *
