"""`from preprocessing.readfiles import *` -> the loaders (reference preprocessing/readfiles.py)."""
from vehicles_dispatch_simulator_b200.readfiles import *  # noqa: F401,F403
