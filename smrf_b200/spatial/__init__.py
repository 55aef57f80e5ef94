from .idw import IDW  # noqa: F401
from .grid import GRID  # noqa: F401
