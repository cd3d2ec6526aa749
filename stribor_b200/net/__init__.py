from .mlp import *
from .time_net import *
