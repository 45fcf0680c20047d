"""cogsworth_b200 -- B200-native (sm_100a) galactic-orbit integration for cogsworth populations.

Drop-in for the one hot path of TomWagg/cogsworth:
``Population.perform_galactic_evolution`` -> ``kicks.integrate_orbit_with_events`` ->
gala ``integrate_orbit`` (Leapfrog / DOPRI853 in MilkyWayPotential v1 / v2) with supernova kicks.
Everything numerical runs in hand-written CUDA kernels behind the C ABI in
``include/cogsworth_b200.h``; there is no CPU fallback.
"""
from . import constants, potential  # noqa: F401
from .potential import CompositePotential, MilkyWayPotential, MilkyWayPotential2022  # noqa: F401
from .batch import EventTable, OrbitBatch, build_orbit_batch  # noqa: F401
from .orbits import Orbit, OrbitArray  # noqa: F401
from .orbit_io import load_orbits, save_orbits  # noqa: F401
from .events import identify_events, identify_event_tables  # noqa: F401
from .kicks import get_kick_differential, integrate_orbit_with_events, integrate_orbits  # noqa: F401
from .rewind import rewind_orbits  # noqa: F401
from .classify import get_rel_speed  # noqa: F401
from .pop import B200Population, GalacticStageB200, InitialGalaxy, TablePopulation  # noqa: F401
from ._lib import Context, NativeLibraryError, default_context  # noqa: F401

__version__ = "0.1.0"
