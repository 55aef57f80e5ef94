from .idw import IDW  # noqa: F401
from .grid import GRID  # noqa: F401
from .dk import DK  # noqa: F401
from . import detrended_kriging  # noqa: F401
