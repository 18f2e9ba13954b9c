"""Import alias: the product package lives in the directory `itensorqtt.jl_b200/` (the name the build
contract fixes), which is not a valid Python identifier, so this shim re-points the package path."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "itensorqtt.jl_b200"))

from ._lib import *        # noqa: F401,F403,E402
from .chains import *      # noqa: F401,F403,E402
from .operators import *   # noqa: F401,F403,E402
from .api import *         # noqa: F401,F403,E402
from .batch import *     # noqa: F401,F403,E402
