"""Cluster / Grid / Vehicle / Order / Transition -- the reference's object API
(objects/objects.py:6-168): same class names, positional constructor
signatures and attribute names, so agent code that walks
`sim.Clusters[i].IdleVehicles` etc. runs unchanged.

In this framework the objects are host-side VIEWS: the authoritative state is
the SoA tensors in HBM (engine.DispatchEngine); Simulation refreshes these
objects from the device before it calls an overridden hook and reads the
agent's list/dict mutations back afterwards (see simulation.py)."""


class Cluster(object):
    def __init__(self, ID, Nodes, Neighbor, RebalanceNumber, IdleVehicles, VehiclesArrivetime, Orders):
        self.ID = ID
        self.Nodes = Nodes
        self.Neighbor = Neighbor
        self.RebalanceNumber = RebalanceNumber
        self.IdleVehicles = IdleVehicles
        self.VehiclesArrivetime = VehiclesArrivetime
        self.Orders = Orders
        self.PerRebalanceIdleVehicles = 0
        self.LaterRebalanceIdleVehicles = 0
        self.PerMatchIdleVehicles = 0
        self.RebalanceFrequency = 0

    def Reset(self):
        self.RebalanceNumber = 0
        self.IdleVehicles.clear()
        self.VehiclesArrivetime.clear()
        self.Orders.clear()
        self.PerRebalanceIdleVehicles = 0
        self.PerMatchIdleVehicles = 0

    def ArriveClusterUpDate(self, vehicle):
        self.IdleVehicles.append(vehicle)
        self.VehiclesArrivetime.pop(vehicle)

    def Example(self):
        print("Order Example output")
        for k in ("ID", "Nodes", "Neighbor", "RebalanceNumber", "IdleVehicles", "VehiclesArrivetime", "Orders"):
            print(k + ":", getattr(self, k))


class Grid(Cluster):
    """Same fields as Cluster (objects/objects.py:131-168); Example prints
    neighbour IDs."""

    def __init__(self, ID, Nodes, Neighbor, RebalanceNumber, IdleVehicles, VehiclesArrivetime, Orders):
        Cluster.__init__(self, ID, Nodes, Neighbor, RebalanceNumber, IdleVehicles, VehiclesArrivetime, Orders)
        del self.RebalanceFrequency

    def Example(self):
        print("ID:", self.ID)
        print("Nodes:", self.Nodes)
        print("Neighbor:[", end=' ')
        for i in self.Neighbor:
            print(i.ID, end=' ')
        print("]")
        print("RebalanceNumber:", self.RebalanceNumber)
        print("IdleVehicles:", self.IdleVehicles)
        print("VehiclesArrivetime:", self.VehiclesArrivetime)
        print("Orders:", self.Orders)
        print()


class Order(object):
    def __init__(self, ID, ReleasTime, PickupPoint, DeliveryPoint, PickupTimeWindow, PickupWaitTime, ArriveInfo, OrderValue):
        self.ID = ID
        self.ReleasTime = ReleasTime
        self.PickupPoint = PickupPoint
        self.DeliveryPoint = DeliveryPoint
        self.PickupTimeWindow = PickupTimeWindow
        self.PickupWaitTime = PickupWaitTime
        self.ArriveInfo = ArriveInfo
        self.OrderValue = OrderValue

    def ArriveOrderTimeRecord(self, ArriveTime):
        self.ArriveInfo = "ArriveTime:" + str(ArriveTime)

    def Example(self):
        print("Order Example output")
        for k in ("ID", "ReleasTime", "PickupPoint", "DeliveryPoint", "PickupTimeWindow", "PickupWaitTime", "ArriveInfo"):
            print(k + ":", getattr(self, k))
        print()

    def Reset(self):
        self.PickupWaitTime = None
        self.ArriveInfo = None


class Vehicle(object):
    def __init__(self, ID, LocationNode, Cluster, Orders, DeliveryPoint):
        self.ID = ID
        self.LocationNode = LocationNode
        self.Cluster = Cluster
        self.Orders = Orders
        self.DeliveryPoint = DeliveryPoint

    def ArriveVehicleUpDate(self, DeliveryCluster):
        self.LocationNode = self.DeliveryPoint
        self.DeliveryPoint = None
        self.Cluster = DeliveryCluster
        if len(self.Orders):
            self.Orders.clear()

    def Reset(self):
        self.Orders.clear()
        self.DeliveryPoint = None

    def Example(self):
        print("Vehicle Example output")
        for k in ("ID", "LocationNode", "Cluster", "Orders", "DeliveryPoint"):
            print(k + ":", getattr(self, k))
        print()


class Transition(object):
    """RL tuple; the reference never instantiates it (objects/objects.py:105-128)."""

    def __init__(self, FromCluster, ArriveCluster, Vehicle, State, StateQTable, Action, TotallyReward,
                 PositiveReward, NegativeReward, NeighborNegativeReward, State_, State_QTable):
        self.FromCluster = FromCluster
        self.ArriveCluster = ArriveCluster
        self.Vehicle = Vehicle
        self.State = State
        self.StateQTable = StateQTable
        self.Action = Action
        self.TotallyReward = TotallyReward
        self.PositiveReward = PositiveReward
        self.NegativeReward = NegativeReward
        self.NeighborNegativeReward = NeighborNegativeReward
        self.State_ = State_
        self.State_QTable = State_QTable

    def Example(self):
        print("Transition Example output")
        for k in ("Action", "TotallyReward", "PositiveReward", "NegativeReward", "NeighborNegativeReward"):
            print(k + ":", getattr(self, k))
        print()
