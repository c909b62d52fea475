"""`from objects.objects import ...` -> the B200 object views (reference objects/objects.py)."""
from vehicles_dispatch_simulator_b200.objects import Cluster, Grid, Order, Transition, Vehicle  # noqa: F401
from vehicles_dispatch_simulator_b200.setting import *  # noqa: F401,F403
