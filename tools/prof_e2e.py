"""Where does the host-side time of one public call go?  python tools/prof_e2e.py [R]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isingmontecarlo_b200 import MODE_FAST, lattices  # noqa: E402
from isingmontecarlo_b200.sse import QmcIsingGraph  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = QmcIsingGraph(lattices.square_periodic(32, -1.0), 3.04, 0.0, 1024, 0x55E00000 + np.arange(R, dtype=np.uint64), 16.0, mode=MODE_FAST)
g.timesteps(60, 16.0)
betas = np.full(R, 16.0)


def timeit(name, fn, n=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    print(f"{name:45s} {(time.perf_counter() - t0) / n * 1e3:8.2f} ms")


timeit("enqueue_sweeps(1) + synchronize", lambda: (g.enqueue_sweeps(1), g.synchronize()))
timeit("timesteps(1)  [energies only]", lambda: g.timesteps(1))
timeit("timesteps_sample(1, None, 1)", lambda: g.timesteps_sample(1, None, 1))
def forced():
    g._betas = None
    return g.timesteps_sample(1, betas, 1)
timeit("timesteps_sample(1, betas, 1) forced set_betas", forced)
timeit("state_ref()", lambda: g.state_ref())
timeit("get_n()", lambda: g.get_n())
