"""CPU: pins the oracle (oracle/vds_oracle.c) to traces recorded from the
UNMODIFIED Python reference -- per-order vehicle / wait, per-tick per-cluster
PerMatch / PerDispatch / LaterDispatch idle counts, SupplyExpect, the ORDER of
every Cluster.IdleVehicles list, counters -- on the committed small-city
fixtures and (when tests/golden/_real/ is present) the three full-day
shipped-data configurations of SURVEY 8c."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN, REAL_CASES, SMALL_CASES, check_oracle_against_golden, load_golden


@pytest.mark.parametrize("name", SMALL_CASES)
def test_oracle_small_city(name):
    assert check_oracle_against_golden(load_golden(name)) == 148


GOLDEN_FINAL = {   # SURVEY.md 8c / BASELINE.md table
    "kmeans": (209422, 141716, 102637, 1671284),
    "grid6000": (209422, 72496, 145514, 3375797),
    "grid5000d3": (209422, 46916, 405749, 4003693),
    # SURVEY 8d's parity variant of BASELINE config 5: real city, Grid, SideLengthMeter=400 (32 x 24 cells, 111 empty), V=10000
    "grid400v10000": (209422, 81169, 55940, 3167227),
}


@pytest.mark.parametrize("name", REAL_CASES)
def test_oracle_real_day(name):
    z = load_golden(name, real=True)
    if z is None:
        pytest.skip("tests/golden/_real not present (regenerate with make_golden.py --real)")
    assert tuple(int(x) for x in z["tr_final"][:4]) == GOLDEN_FINAL[name]
    assert check_oracle_against_golden(z) == 148
    dg = json.load(open(os.path.join(GOLDEN, "real_digest.json")))[name]
    h = hashlib.sha256(np.ascontiguousarray(z["tr_order_vehicle"].astype(np.int32)).tobytes()).hexdigest()
    assert h == dg["order_vehicle_sha256"]


def test_digest_matches_survey_table():
    dg = json.load(open(os.path.join(GOLDEN, "real_digest.json")))
    for name, exp in GOLDEN_FINAL.items():
        assert tuple(dg[name]["final"][:4]) == exp
        assert dg[name]["final"][6] == 148
