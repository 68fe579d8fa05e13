"""north_star's statistical criterion on the CUDA path: energy, magnetisation and susceptibility estimators agree
with exact diagonalisation within 3 sigma of the statistical error over independent chains (E per
chain from S7; m = (1/N) sum (2 s_i - 1) from state_ref samples at p = 0; chi = beta N (<m^2> - <|m|>^2); errors
by binning over chains, not time) -- in both cluster orders and with both diagonal rules."""
import numpy as np
import pytest

from isingmontecarlo_b200 import MODE_FAST, MODE_STRICT, lattices
from tests.ed import tfim_thermal

pytestmark = pytest.mark.gpu

SYSTEMS = [
    ("ring6_ferro", lattices.one_d_periodic(6, -1.0), 0.7, 0.0, 3.0),
    ("frustrated_h", [((0, 1), 1.0), ((1, 2), 1.0), ((2, 0), 1.0), ((2, 3), 1.0), ((3, 0), 1.0)], 0.5, -0.3, 2.0),
    ("square3_mixed", lattices.two_d_periodic_mixed(3), 1.0, 0.0, 1.5),
]


@pytest.mark.parametrize("mode,heatbath", [(MODE_STRICT, False), (MODE_FAST, False), (MODE_FAST, True)])
@pytest.mark.parametrize("name,edges,gamma,h,beta", SYSTEMS)
def test_estimators_match_exact_diagonalisation(name, edges, gamma, h, beta, mode, heatbath):
    from isingmontecarlo_b200.sse import QmcIsingGraph

    nv = lattices.nvars_of(edges)
    exact = tfim_thermal(edges, nv, gamma, h, beta)
    chains = 256
    g = QmcIsingGraph(edges, gamma, h, nv, 0xE57A0000 + 4096 * mode + 1024 * heatbath + np.arange(chains, dtype=np.uint64), beta, mode=mode)
    g.set_enable_heatbath(heatbath)
    g.timesteps(500, beta)
    samples, e = g.timesteps_sample(4000, beta, 4)
    assert g.verify(0) and g.verify(chains - 1)
    m = (2.0 * samples.astype(np.float64) - 1.0).mean(axis=2)  # [chain][sample]
    est = {"E": e, "m": m.mean(axis=1), "m2": (m * m).mean(axis=1), "absm": np.abs(m).mean(axis=1)}
    for k, per_chain in est.items():
        mean, err = per_chain.mean(), per_chain.std(ddof=1) / np.sqrt(chains)
        assert abs(mean - exact[k]) < 3.0 * err + 1e-9, (name, k, mean, err, exact[k])
    # susceptibility: jackknife over chains
    m2c, amc = est["m2"], est["absm"]
    chi = beta * nv * (m2c.mean() - amc.mean() ** 2)
    jk = np.array([beta * nv * (np.delete(m2c, i).mean() - np.delete(amc, i).mean() ** 2) for i in range(chains)])
    chi_err = np.sqrt((chains - 1) * ((jk - jk.mean()) ** 2).mean())
    chi_exact = beta * nv * (exact["m2"] - exact["absm"] ** 2)
    assert abs(chi - chi_exact) < 3.0 * chi_err + 1e-9, (name, chi, chi_err, chi_exact)
