"""RVB update of the oracle (oracle.c::orc_sse_rvb_update, restating rvb.rs:60-1221 and util/bondcontainer.rs).  The
reference holds no golden vectors for it (its RVB tests are crash tests: tests/check_rvb_crash.rs, the `rvb` cases of
tests/longitudinal_crash.rs), so the restatement is pinned the way the rest of the oracle is: the reference's own crash
tests with its invariant verify(), and exact diagonalisation of frustrated models -- a wrong acceptance ratio or a wrong
rotation breaks detailed balance and shows up in the energy."""
import numpy as np
import pytest

from isingmontecarlo_b200 import lattices
from oracle import pyoracle as po
from tests.ed import tfim_thermal

TRI2 = [((0, 1), 1.0), ((1, 2), 1.0), ((2, 0), 1.0), ((2, 3), 1.0), ((3, 4), 1.0), ((4, 2), 1.0)]
DIAMOND = [((0, 1), 1.0), ((1, 2), 1.0), ((2, 0), 1.0), ((2, 3), 1.0), ((3, 0), 1.0)]


@pytest.mark.parametrize("edges,gamma,h,nvars", [
    (lattices.two_d_periodic_mixed(3), 0.1, 0.0, 9),   # check_rvb_crash.rs:296-315 run_three
    (lattices.two_d_periodic_mixed(4), 0.1, 0.0, 16),  # :318-337 run_four
    (lattices.two_unit_cell(), 1.0, 0.0, 8),           # :340-359 run_two_unit_cell
    (lattices.two_d_periodic_mixed(3), 1.0, 1.0, 9),   # longitudinal_crash.rs rvb cases: h != 0 takes rvb_update_with_ising_weight
    (lattices.two_unit_cell(), 1.0, 1.0, 8),
])
@pytest.mark.parametrize("mode", [po.MODE_STRICT, po.MODE_FAST, po.MODE_COUNTER])
def test_rvb_crash_tests_keep_the_invariant(edges, gamma, h, nvars, mode):
    for seed in range(4):
        q = po.SseOracle(edges, gamma, h, nvars, key=seed, state=[0] * nvars)
        q.set_run_rvb(True)
        for _ in range(250):
            q.timestep(1.0, mode)
            assert q.error == 0
            assert q.verify()
        assert 0.0 < q.rvb_success_rate() < 1.0


def test_single_rvb_sweep_counts_and_keeps_the_invariant():
    q = po.SseOracle(lattices.two_unit_cell(), 1.0, 0.0, 8, key=3)
    q.timesteps(50, 2.0)
    n = q.n
    succ, att = q.single_rvb_sweep()
    assert att == (8 + 1) // 2 and 0 <= succ <= att  # qmc_ising.rs:375
    succ, att = q.single_rvb_sweep(40)
    assert att == 40 and 0 < succ <= 40
    assert q.error == 0 and q.verify() and q.n == n  # the update moves and flips ops, it never adds or removes one


@pytest.mark.parametrize("name,edges,gamma,h,beta", [
    ("two_triangles", TRI2, 0.3, 0.0, 3.0),
    ("diamond_mixed_J", [((0, 1), 1.0), ((1, 2), 0.7), ((2, 0), 1.3), ((2, 3), -1.0), ((3, 0), 1.0)], 0.4, 0.0, 2.5),
    ("diamond_h", DIAMOND, 0.4, 0.25, 2.0),
    ("triangular_torus_3x4", [((jj * 3 + i, k), 1.0) for i in range(3) for jj in range(4)
                              for k in (jj * 3 + (i + 1) % 3, ((jj + 1) % 4) * 3 + i, ((jj + 1) % 4) * 3 + (i + 1) % 3)], 0.7, 0.0, 3.0),
])
def test_rvb_energy_matches_exact_diagonalisation(name, edges, gamma, h, beta):
    nvars = lattices.nvars_of(edges)
    exact = tfim_thermal(edges, nvars, gamma, h, beta)
    chains = 32
    reps = [po.SseOracle(edges, gamma, h, nvars, key=0x7E0 + r) for r in range(chains)]
    for r in reps:
        r.set_run_rvb(True)
    po.sse_batch_timesteps(reps, 500, [beta] * chains, po.MODE_STRICT)
    _, e = po.sse_batch_timesteps(reps, 8000, [beta] * chains, po.MODE_STRICT)
    assert all(r.error == 0 and r.verify() for r in reps)
    assert all(r.rvb_success_rate() > 0.05 for r in reps)  # the move is doing work in these models
    mean, err = e.mean(), e.std(ddof=1) / np.sqrt(chains)
    assert abs(mean - exact["E"]) < 3.0 * err + 1e-9, (name, mean, err, exact["E"])


def _op(bond, ins, outs):
    return bond | (ins << 24) | (outs << 26)


def test_hand_derived_single_variable_cluster_flip():
    """The set-up of tests/check_rvb_crash.rs:69-109 (one variable, two constant ops, no edges), one proposal under a
    scripted stream, derived by hand from rvb.rs:
    word 0 -> gen_range(0..2) = 0: the constant op at p = 0 (var 0, flip index 0)                       (:125-139)
    word 0 -> contiguous_bits = 0: cluster size 1                                                        (:148, :1190)
    build_cluster: push (0, Some(0)) weight 1; pop: f_ratio = 1/1 -> gen_bool(1.0) draws nothing; get_random draws one
      word and takes the only key; its neighbours in time, flip_dec = flip_inc = 1, go to the boundary    (:1054-1090)
    toggles = [p(0), p(1)] = [0, 1]; no bonds -> p_to_flip = 1.0 -> accepted without a draw               (:182-203, :247-252)
    mutate_graph: the world-line piece between the two ops flips: op 0 becomes 0 -> 1, op 1 becomes 1 -> 0; the state at
      p = 0 is outside the piece and stays                                                               (:434-469)"""
    g = po.SseOracle([], 1.0, 0.0, 2, key=0, state=[0], nvars=1)
    g.load_ops([_op(0, 0, 0), _op(0, 0, 0)], [0])
    g.set_script([0, 0, 12345])
    assert g.rvb_update(1) == 1
    assert g.error == 0 and g.cursor == 3
    assert list(g.dump_ops()) == [_op(0, 0, 1), _op(0, 1, 0)]
    assert list(g.state()) == [0] and g.verify()


def test_hand_derived_rotation_of_a_border_op():
    """Antiferromagnetic triangle, bonds b0 = (0,1), b1 = (1,2), b2 = (2,0), J = 1, no transverse ops in the string; state
    (0,1,0); one diagonal op on b1 with spins (1,0).  One proposal, by hand from rvb.rs:
    word 0xC000.. -> gen_range(0..3) = 2: no constant ops anywhere, so the third op-less variable, var 2, flip None  (:125-146)
    word 0 -> cluster size 1.  build_cluster: push (2, None); pop: f_ratio = 0/1 -> gen_bool(0.0) DRAWS a word and says no;
      get_random draws a word, takes var 2; neighbours 1 (via b1) and 0 (via b2) go to the boundary with |J|          (:1054-1118)
    sub-variables {0,1,2}, the whole world line of var 2 is the cluster (starting state set)                          (:182-203)
    calculate_flip_prob: border before = {b1: 2, b2: 0}, after the flip {b1: 0, b2: 2}; the op on b1 counts n = 1;
      totals are equal -> multiplier 1.0 -> accepted without a draw                                                   (:617-646, :843-851, :1194-1221)
    mutate_graph: substate with the cluster flipped = (0,1,1); border bonds with their new weights, in insertion order
      [b1: 0, b2: 2]; the op sits on a border bond, so it is rotated: get_random draws gen_range(0. ..2.0) = 1.0 from
      word 2^63, passes b1 (weight 0) and lands on b2 = (2,0) with spins (1,0)                                        (:366-381, :411-432)
    the cluster reaches p = 0, so state[2] flips                                                                      (:262-276)"""
    tri = [((0, 1), 1.0), ((1, 2), 1.0), ((2, 0), 1.0)]
    g = po.SseOracle(tri, 1.0, 0.0, 1, key=0, state=[0, 1, 0], nvars=3)
    g.load_ops([_op(1, 0b01, 0b01)], [0, 1, 0])  # bond 1 = (1,2): first variable (1) up, second (2) down
    g.set_script([0xC000000000000000, 0, 777, 888, 1 << 63])
    assert g.rvb_update(1) == 1
    assert g.error == 0 and g.cursor == 5
    assert list(g.dump_ops()) == [_op(2, 0b01, 0b01)]  # bond 2 = (2,0): variable 2 up, variable 0 down
    assert list(g.state()) == [0, 1, 1] and g.verify()
