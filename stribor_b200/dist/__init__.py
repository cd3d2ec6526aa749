from .normal import *
