"""alias package, see vehicles_dispatch_simulator_b200/compat/__init__.py"""
