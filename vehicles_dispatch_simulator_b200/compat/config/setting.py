"""`from config.setting import *` -> the same constants (reference config/setting.py)."""
from vehicles_dispatch_simulator_b200.setting import *  # noqa: F401,F403
