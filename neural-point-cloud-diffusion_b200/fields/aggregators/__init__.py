from .aggregator import Aggregator
from .mlp import MLP
