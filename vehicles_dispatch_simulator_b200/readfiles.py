"""CSV loaders with the reference's names and return shapes
(preprocessing/readfiles.py:1-127).  One-time host I/O, not on the hot path.

ReadOrder keeps pandas' default (unstable) sort_values on the minute-truncated
Start_time: the intra-minute order -- and therefore Order.ID and the processing
order -- is whatever this pandas build produces, exactly as in the reference
(SURVEY Q9)."""
import datetime as dt
import os
from datetime import datetime

import numpy as np
import pandas as pd


def timestamp_datetime(value):
    d = datetime.fromtimestamp(value)                  # local TZ, as the reference
    return dt.datetime(d.year, d.month, d.day, d.hour, d.minute, 0)


def string_datetime(value):
    return dt.datetime.strptime(value, "%Y-%m-%d %H:%M:%S")


def string_pdTimestamp(value):
    d = string_datetime(value)
    return pd.Timestamp(d.year, d.month, d.day, d.hour, d.minute)


def ReadCostMap(input_file_path):
    """DataFrame (NOT .values): Simulation.RoadCost indexes it column-first,
    which is what makes RoadCost(s, e) == int(A[e, s])."""
    return pd.read_csv(input_file_path, header=None)


def ReadNode(input_file_path):
    return pd.read_csv(input_file_path)


def ReadNodeIDList(input_file_path):
    with open(input_file_path, "r") as f:
        return [int(line.split()[0]) for line in f if line.strip()]


def ReadOrder(input_file_path):
    Order = pd.read_csv(input_file_path)
    Order = Order.drop(columns=['End_time', 'PointS_Longitude', 'PointS_Latitude', 'PointE_Longitude', 'PointE_Latitude'])
    Order["Start_time"] = Order["Start_time"].apply(timestamp_datetime)
    Order = Order.sort_values(by="Start_time")
    Order["ID"] = range(0, Order.shape[0])
    return Order.values


def ReadResetOrder(input_file_path):
    return pd.read_csv(input_file_path).values


def ReadDriver(input_file_path="./data/Drivers1101.csv"):
    Driver = pd.read_csv(input_file_path)
    return Driver.drop(columns=['Start_time']).values


def ReadAllFiles(OrderFileDate="1101", data_dir=None):
    d = data_dir or os.path.join(os.getcwd(), "data")
    Node = ReadNode(os.path.join(d, "Node.csv"))
    NodeIDList = ReadNodeIDList(os.path.join(d, "NodeIDList.txt"))
    Orders = ReadOrder(os.path.join(d, "order_2016" + OrderFileDate + ".csv"))
    Vehicles = ReadDriver(os.path.join(d, "Drivers1101.csv"))
    Map = ReadCostMap(os.path.join(d, "AccurateMap.csv"))
    return Node, NodeIDList, Orders, Vehicles, Map


def ReadOrdersVehiclesFiles(OrderFileDate="1101", data_dir=None):
    d = data_dir or os.path.join(os.getcwd(), "data")
    return (ReadOrder(os.path.join(d, "order_2016" + OrderFileDate + ".csv")),
            ReadDriver(os.path.join(d, "Drivers1101.csv")))
