"""CPU oracle of the galactic-orbit hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  cogsworth_b200 (the product) never does.
"""
from .oracle import *  # noqa: F401,F403
