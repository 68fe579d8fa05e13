"""isingmontecarlo_b200 -- B200-native replica engine for the data-parallel hot path of
Renmusxd/IsingMonteCarlo (SSE transverse-field Ising sweeps, classical checkerboard sweeps,
parallel-tempering swaps).  The compute path is the CUDA library behind include/qmcb.h; this
package is the thin host-side mirror of the reference's interface for that path."""
from . import lattices  # noqa: F401
from ._lib import MODE_COUNTER, MODE_FAST, MODE_STRICT, OP_EMPTY, QmcbError  # noqa: F401


def __getattr__(name):
    if name in ("QmcIsingGraph", "DefaultQmcIsingGraph"):
        from .sse import QmcIsingGraph
        return QmcIsingGraph
    if name == "GraphState":
        from .classical import GraphState
        return GraphState
    if name == "TemperingContainer":
        from .tempering import TemperingContainer
        return TemperingContainer
    raise AttributeError(name)
