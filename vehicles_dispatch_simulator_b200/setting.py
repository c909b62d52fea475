"""Simulator constants -- same names and values as the reference's
config/setting.py:1-31 so `from ...setting import *` keeps working."""
import numpy as np

MINUTES = 60000000000
TIMESTEP = np.timedelta64(10 * MINUTES)
PICKUPTIMEWINDOW = np.timedelta64(10 * MINUTES)

NeighborCanServer = False
FocusOnLocalRegion = False
LocalRegionBound = (104.035, 104.105, 30.625, 30.695)
if FocusOnLocalRegion == False:  # noqa: E712  (kept as in the reference)
    LocalRegionBound = (104.011, 104.125, 30.618, 30.703)

VehiclesNumber = 6000
SideLengthMeter = 800
VehiclesServiceMeter = 800

DispatchMode = "Simulation"
DemandPredictionMode = "None"
# ["TransportationClustering","KmeansClustering","SpectralClustering"]
ClusterMode = "Grid"
