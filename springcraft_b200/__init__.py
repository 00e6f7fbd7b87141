"""springcraft_b200 -- B200-native (sm_100a) elastic-network hot path with the
springcraft API surface (flat namespace like springcraft/__init__.py:12-15)."""

__version__ = "0.3.0+b200.1"
__reference_version__ = "0.3.0"

from . import nma  # noqa: F401
from .anm import *  # noqa: F401,F403
from .dense_solver import *  # noqa: F401,F403
from .ensemble import *  # noqa: F401,F403
from .forcefield import *  # noqa: F401,F403
from .gnm import *  # noqa: F401,F403
from .interaction import *  # noqa: F401,F403
from .structure import (AtomArray, BadStructureError, read_cif_ca, read_pdb_ca,  # noqa: F401
                        read_pdb_ca_models)
