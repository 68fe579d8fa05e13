"""Small runs for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, tiny sizes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isingmontecarlo_b200 import MODE_COUNTER, MODE_FAST, MODE_STRICT, lattices  # noqa: E402
from isingmontecarlo_b200.classical import GraphState  # noqa: E402
from isingmontecarlo_b200.sse import QmcIsingGraph  # noqa: E402

edges = lattices.two_d_periodic_mixed(4)
for mode in (MODE_FAST, MODE_STRICT, MODE_COUNTER):
    for hb in (False, True):
        if hb and mode == MODE_COUNTER:
            continue
        for pipeline in (1, 0):
            g = QmcIsingGraph(edges, 1.0, 0.4, 16, [1, 2, 3, 4, 5], 1.5, mode=mode)
            g.set_option("pipeline", pipeline)
            g.set_enable_heatbath(hb)
            g.timesteps(12, 1.5)
            assert g.verify()
            g.imaginary_time_magnetization()
            g.close()
g = QmcIsingGraph(lattices.square_periodic(8, -1.0), 3.04, 0.0, 64, np.arange(40, dtype=np.uint64) + 7, 2.0, mode=MODE_FAST)
g.set_option("minblocks", 7)
g.timesteps(8, 2.0)
g.calculate_variable_autocorrelation(16, 2.0, 1)
g.set_option("minblocks", 0)
g.close()
for layout in (0, 1, 7):
    g = QmcIsingGraph(lattices.square_periodic(8, -1.0), 3.04, 0.0, 64, np.arange(9, dtype=np.uint64) + 7, 2.0, mode=MODE_STRICT)
    g.set_option("strict_layout", layout)
    g.timesteps(6, 2.0)
    assert g.verify()
    g.close()
g = QmcIsingGraph(lattices.square_periodic(8, -1.0), 3.04, 0.0, 64, np.arange(40, dtype=np.uint64) + 7, 2.0, mode=MODE_COUNTER)
g.timesteps(8, 2.0)
assert g.verify()
g.close()
c = GraphState(lattices.square_periodic(64, -1.0), np.zeros(64 * 64), [1, 2], 0.44)
c.sweeps(3)
c.get_energy()
c.do_time_step(20, 10, 2)
c.close()
c = GraphState(lattices.triangular_periodic(6, 1.0), np.zeros(36), [1, 2, 3], 0.7)
c.sweeps(3)
for _ in range(6):
    c.do_time_step()
c.close()
print("sanitize run ok")
