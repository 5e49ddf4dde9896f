"""uvic2.9_b200 -- B200-native ocean tracer step for the UVic ESCM 2.9 (load as uvic29_b200)."""
from . import synthetic  # noqa: F401
from . import timestep  # noqa: F401
from . import api  # noqa: F401
from .api import TracerContext, TracerGroup, UvicError, load_library  # noqa: F401
from . import slab  # noqa: F401
from . import mobi_params  # noqa: F401
