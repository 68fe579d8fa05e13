"""Hand-derived known answers for the reference algorithm, run against
the CPU oracle with a scripted word stream.  The strings B1-B3 are the ones built by hand in the
reference's tests/cluster_test.rs:7-75."""
import numpy as np

from oracle import pyoracle as po

FLIP, KEEP = 0, 1 << 63  # gen_bool(0.5) is `word < 2^63`


def word(bond, ins, outs):
    w = bond
    for r, (i, o) in enumerate(zip(ins, outs)):
        w |= int(i) << (24 + r)
        w |= int(o) << (26 + r)
    return w


def make(nvars, edges, ops, state, transverse=1.0, longitudinal=0.0):
    g = po.SseOracle(edges, transverse, longitudinal, len(ops), key=0, state=state, nvars=nvars)
    g.load_ops(ops, state)
    return g


def test_b1_single_site_op():
    # cluster_test.rs:7-21
    g = make(1, [], [word(0, [0], [0])], [0])
    g.set_script([FLIP])
    assert g.single_cluster_step() == 1
    bi, bo = g.boundaries(1)
    assert (bi[0], bo[0]) == (0, 0)
    assert list(g.state()) == [1]
    assert list(g.dump_ops()) == [word(0, [1], [1])]
    assert g.cursor == 1 and g.verify()


def test_b2_two_site_ops():
    # cluster_test.rs:24-44
    g = make(1, [], [word(0, [0], [0]), word(0, [0], [0])], [0])
    g.set_script([FLIP, KEEP])
    assert g.single_cluster_step() == 2
    bi, bo = g.boundaries(2)
    assert list(zip(bi, bo)) == [(0, 1), (1, 0)]
    assert list(g.state()) == [1]
    assert list(g.dump_ops()) == [word(0, [1], [0]), word(0, [0], [1])]
    assert g.verify()


def test_b3_two_variables():
    # cluster_test.rs:47-75
    ops = [word(0, [0], [0]), word(0, [0], [0]), word(1, [0], [0]), word(1, [0], [0])]
    g = make(2, [], ops, [0, 0])
    g.set_script([KEEP, KEEP, FLIP, KEEP])
    assert g.single_cluster_step() == 4
    bi, bo = g.boundaries(4)
    assert list(zip(bi, bo)) == [(0, 1), (1, 0), (2, 3), (3, 2)]
    assert list(g.state()) == [0, 1]
    assert g.verify()


def test_b4_bond_op_is_interior():
    edges = [((0, 1), -1.0)]
    ops = [word(1, [0], [0]), word(0, [0, 0], [0, 0]), word(2, [0], [0])]
    g = make(2, edges, ops, [0, 0])
    g.set_script([FLIP])
    assert g.single_cluster_step() == 1
    bi, bo = g.boundaries(3)
    assert list(zip(bi, bo)) == [(0, 0)] * 3
    assert list(g.state()) == [1, 1]
    assert list(g.dump_ops()) == [word(1, [1], [1]), word(0, [1, 1], [1, 1]), word(2, [1], [1])]
    assert g.verify()


def test_b5_no_edge_op_is_one_cluster():
    # cluster.rs:98-107: even a disconnected bond graph is ONE cluster
    edges = [((0, 1), -1.0), ((2, 3), -1.0)]
    ops = [word(0, [0, 0], [0, 0]), word(1, [1, 1], [1, 1])]
    for mode in (po.MODE_STRICT, po.MODE_FAST):
        g = make(4, edges, ops, [0, 0, 1, 1])
        g.set_script([FLIP])
        assert g.single_cluster_step(mode) == 1
        assert g.verify()


def test_d1_diagonal_rules_and_full_step():
    g = po.SseOracle([((0, 1), 1.0)], 1.0, 0.0, 2, key=0, state=[0, 1])
    g.set_script([0, 1 << 63, FLIP])
    g.timestep(1.0)
    assert g.error == 0
    assert g.n == 2 and g.cursor == 3 and g.cutoff == 3
    assert list(g.state()) == [1, 0]
    ops = g.dump_ops()
    assert list(ops) == [word(0, [1, 0], [1, 0]), word(1, [1], [1]), po.OP_EMPTY]
    assert g.verify()
    # next sweep, p0 is a diagonal bond op: num = 6, den = (3-2)+1 = 2 -> one draw, removed iff
    # word < 0x5555555555555400; p1 (transverse, num=3, den = 2+1): den == num -> gen_bool(1.0),
    # removed with no draw; p2 empty.
    g.set_script([0x5555555555555400, 0, 0, 0])
    g.set_cursor(0)
    g.single_diagonal_step(1.0)
    # p0 kept (word not below threshold); p1: den = (3-2)+1 = 2 < num = 3 -> draw word 0 < thr -> removed
    assert g.n >= 1


def test_d2_d3_draw_counts():
    g = po.SseOracle([((0, 1), 1.0)], 1.0, 0.0, 3, key=0, state=[0, 0])
    # p0: b=1 (transverse) num=3 den=3 -> gen_bool(1.0): accepted, NO draw
    # p1: b=0 (aligned, J>0 -> weight 0) -> gen_bool(0.0) consumes one draw, rejected
    # p2: b=1 num=3 > den=2 -> accepted, no draw
    g.set_script([1 << 63, 0, 12345, 1 << 63])
    g.single_diagonal_step(1.0)
    assert g.error == 0
    assert g.cursor == 4 and g.n == 2
    assert list(g.dump_ops()) == [word(1, [0], [0]), po.OP_EMPTY, word(1, [0], [0])]


def test_fast_mode_labels_are_min_segment_ids():
    # B3 string: four site ops, two per variable -> segments {0: var0 wrap, 1: var1 wrap, 2.., 5}
    ops = [word(0, [0], [0]), word(0, [0], [0]), word(1, [0], [0]), word(1, [0], [0])]
    g = make(2, [], ops, [0, 0])
    assert g.single_cluster_step(po.MODE_FAST) == 4
    bi, bo = g.boundaries(4)
    # var0: seg 0 = before p0 == after p1 (closure with id 3), seg 2 = between p0 and p1
    assert list(zip(bi, bo)) == [(0, 2), (2, 0), (1, 4), (4, 1)]
    assert g.cursor == 1 and g.verify()


def test_h1_heatbath_rules_and_draw_order():
    # heatbath.rs:149-209 by hand.  Edge (0,1) J=+1, Gamma=1: maximum weights [2, 1, 1] -> cumulative
    # [2, 3, 4], total 4 (heatbath.rs:17-35, :130-146).  beta=1, cutoff 3, state [0,1].
    g = po.SseOracle([((0, 1), 1.0)], 1.0, 0.0, 3, key=0, state=[0, 1])
    g.set_enable_heatbath(True)
    top = (1 << 64) - 1
    g.set_script([
        0,            # p0 empty, n=0: gen_bool(4/7) -> attempt
        1 << 63,      #   p = gen_range(0. ..1.0) = 0.5
        1 << 62,      #   c = 0.25 * 4 = 1.0 -> bond 0 (cum 2 > 1), maxweight 2, anti-aligned weight 2: 1.0 < 2 -> insert
        top,          # p1 empty, n=1: gen_bool(4/6) false -> one draw only
        0,            # p2 empty: attempt
        3 << 62,      #   p = 0.75
        5 << 61,      #   c = 0.625 * 4 = 2.5 -> index 1 (cum 3 > 2.5 > cum 2): transverse op on var 0, 0.75 * 1 < 1 -> insert
    ])
    g.single_diagonal_step(1.0)
    assert g.error == 0 and g.n == 2 and g.cursor == 7 and g.cutoff == 3
    assert list(g.dump_ops()) == [word(0, [0, 1], [0, 1]), po.OP_EMPTY, word(1, [0], [0])]
    assert g.verify()
    # removal: p0 diagonal, n=2: gen_bool(2/6): 0x5555555555555400 is the threshold itself -> kept;
    # p1 empty: gen_bool(4/5) false; p2 diagonal: gen_bool(2/6) with word 0 -> removed
    g.set_script([0x5555555555555400, top, 0])
    g.set_cursor(0)
    g.single_diagonal_step(1.0)
    assert g.error == 0 and g.n == 1 and g.cursor == 3
    assert list(g.dump_ops()) == [word(0, [0, 1], [0, 1]), po.OP_EMPTY, po.OP_EMPTY]
    # an insertion attempt that fails the weight test still consumes three words (heatbath.rs:166-187):
    # aligned spins on a J>0 bond have weight 0
    g2 = po.SseOracle([((0, 1), 1.0)], 1.0, 0.0, 1, key=0, state=[0, 0])
    g2.set_enable_heatbath(True)
    g2.set_script([0, 0, 0])  # attempt, p = 0, c = 0 -> bond 0, weight 0: 0 < 0 false
    g2.single_diagonal_step(1.0)
    assert g2.error == 0 and g2.n == 0 and g2.cursor == 3
