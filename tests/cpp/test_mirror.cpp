// The reference's own integration tests, re-read against the C++ mirror (include/qmcb.hpp):
//   tests/longitudinal_crash.rs:39-178  (16 seeds x small lattices with a longitudinal field, 1000 steps, verify())
//   examples/small_qmc.rs               (4-site ring, 1000 steps)
//   tests/convert_test.rs:5-7 lattice   (3-ring, 10 steps)
// Batched: the 16 seeds of a reference test are the 16 replicas of one handle.  Both cluster orders run.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "qmcb.hpp"

using qmcb::Edge;

static std::vector<Edge> two_d_periodic(size_t l) {  // tests/longitudinal_crash.rs:5-23
    std::vector<Edge> right, down;
    auto f = [l](size_t i, size_t j) { return j * l + i; };
    for (size_t i = 0; i < l; i++)
        for (size_t j = 0; j < l; j++) {
            right.push_back({{f(i, j), f((i + 1) % l, j)}, -1.0});
            down.push_back({{f(i, j), f(i, (j + 1) % l)}, i % 2 == 0 ? 1.0 : -1.0});
        }
    right.insert(right.end(), down.begin(), down.end());
    return right;
}
static std::vector<Edge> two_unit_cell() {  // :25-37
    return {{{0, 1}, -1.0}, {{1, 2}, 1.0}, {{2, 3}, 1.0}, {{3, 0}, 1.0}, {{1, 7}, 1.0}, {{4, 5}, -1.0}, {{5, 6}, 1.0}, {{6, 7}, 1.0}, {{7, 4}, 1.0}};
}

#define ASSERT(x)                                                        \
    do {                                                                 \
        if (!(x)) {                                                      \
            std::fprintf(stderr, "assert failed %s:%d: %s\n", __FILE__, __LINE__, #x); \
            std::exit(1);                                                \
        }                                                                \
    } while (0)

int main() {
    std::vector<uint64_t> seeds;
    for (uint64_t i = 0; i < 16; i++) seeds.push_back(i);
    for (int mode : {QMCB_MODE_STRICT, QMCB_MODE_FAST, QMCB_MODE_COUNTER}) {
        {  // run_simple :39-55
            std::vector<bool> st(2, false);
            auto ising = qmcb::DefaultQmcIsingGraph::new_with_rng({{{0, 1}, 1.0}}, 1.0, 1.0, 2, seeds, &st, mode);
            ising.timesteps(1000, 1.0);
            ASSERT(ising.verify());
        }
        {  // run_three / four lattices with longitudinal field :77-133
            for (size_t l : {3u, 4u}) {
                auto ising = qmcb::DefaultQmcIsingGraph::new_with_rng(two_d_periodic(l), 1.0, 1.0, l * l, seeds, nullptr, mode);
                ising.timesteps(1000, 1.0);
                ASSERT(ising.verify());
            }
        }
        {  // two unit cells :135-178
            auto ising = qmcb::DefaultQmcIsingGraph::new_with_rng(two_unit_cell(), 1.0, 1.0, 8, seeds, nullptr, mode);
            ising.timesteps(1000, 1.0);
            ASSERT(ising.verify());
        }
        {  // examples/small_qmc.rs
            auto g = qmcb::DefaultQmcIsingGraph::new_with_rng({{{0, 1}, -1.0}, {{1, 2}, 1.0}, {{2, 3}, 1.0}, {{3, 0}, 1.0}}, 1.0, 0.0, 3, seeds, nullptr, mode);
            auto e = g.timesteps(1000, 1.0);
            double mean = 0;
            for (double x : e) mean += x / e.size();
            ASSERT(mean > -4.6 && mean < -3.6);  // exact -4.055 (DESIGN / SURVEY Appendix E)
            ASSERT(g.verify());
        }
        if (mode != QMCB_MODE_COUNTER) {  // the reference's heat-bath benches (benches/end_to_end.rs:168-258): same lattice, set_enable_heatbath(true)
            auto g = qmcb::DefaultQmcIsingGraph::new_with_rng({{{0, 1}, -1.0}, {{1, 2}, 1.0}, {{2, 3}, 1.0}, {{3, 0}, 1.0}}, 1.0, 0.0, 3, seeds, nullptr, mode);
            g.set_enable_heatbath(true);
            auto e = g.timesteps(1000, 1.0);
            double mean = 0;
            for (double x : e) mean += x / e.size();
            ASSERT(mean > -4.6 && mean < -3.6);
            ASSERT(g.verify());
        }
        {  // tests/check_rvb_crash.rs:340-359 run_two_unit_cell: set_run_rvb(true), timesteps, verify
            auto ising = qmcb::DefaultQmcIsingGraph::new_with_rng({{{0, 1}, -1.0}, {{1, 2}, 1.0}, {{2, 3}, 1.0}, {{3, 0}, 1.0}, {{1, 7}, 1.0}, {{4, 5}, -1.0}, {{5, 6}, 1.0}, {{6, 7}, 1.0}, {{7, 4}, 1.0}},
                                                                  1.0, 0.0, 8, seeds, nullptr, mode);
            ising.set_run_rvb(true);
            ising.timesteps(200, 1.0);
            ASSERT(ising.verify());
            for (double rate : ising.rvb_success_rate()) ASSERT(rate > 0.0 && rate < 1.0);
            auto sw = ising.single_rvb_sweep();
            ASSERT(sw.second == (8 + 1) / 2 && ising.verify());
        }
        {  // convert_test.rs lattice: timestep returns the state, sampling shape
            std::vector<bool> st(3, true);
            auto ising = qmcb::DefaultQmcIsingGraph::new_with_rng({{{0, 1}, 1.0}, {{1, 2}, 1.0}, {{2, 0}, 1.0}}, 1.0, 0.0, 3, {1234}, &st, mode);
            for (int i = 0; i < 10; i++) ASSERT(ising.timestep(1.0)[0].size() == 3);
            auto se = ising.timesteps_sample(20, 1.0, 5);
            ASSERT(se.first[0].size() == 4 && se.first[0][0].size() == 3);
        }
    }
    {  // tempering_container.rs tests (:600-667 run ladders and check verify()): swaps happen, invariants hold
        std::vector<double> betas;
        for (int k = 0; k < 6; k++) betas.push_back(0.4 + 0.3 * k);
        std::vector<uint64_t> keys;
        for (uint64_t s = 0; s < 12; s++) keys.push_back(0x7E00 + s);
        qmcb::TemperingContainer tc(two_d_periodic(4), 1.0, 0.0, 16, betas, 2, keys, 0xBEEF);
        for (int i = 0; i < 30; i++) {
            tc.timesteps(3);
            tc.tempering_step();
        }
        ASSERT(tc.num_graphs() == 12 && tc.get_total_swaps() > 0 && tc.verify());
        auto se = tc.timesteps_sample(12, 2, 3);  // tempering_container.rs:166-208
        ASSERT(se.size() == 12);
        for (auto &slot : se) ASSERT(slot.first.size() == 4 && slot.first[0].size() == 16 && slot.second == slot.second);
        // ParallelTemperingAutocorrelations (:484-630): one autocorrelation per slot, lag 0 is 1
        auto ac = tc.calculate_variable_autocorrelation(64, 2, 2);
        ASSERT(ac.size() == 12 && ac[0].size() == 32);
        for (auto &a : ac) ASSERT(std::fabs(a[0] - 1.0) < 1e-12);
        auto bc = tc.calculate_bond_autocorrelation(64, 2, 2);
        ASSERT(bc.size() == 12 && bc[0].size() == 32 && tc.verify());
    }
    {  // serde round trip (qmc_ising.rs serialize_test): a restored batch continues identically
        auto g = qmcb::DefaultQmcIsingGraph::new_with_rng(two_d_periodic(3), 1.0, 0.3, 9, {21, 22, 23}, nullptr, QMCB_MODE_FAST);
        g.timesteps(50, 1.0);
        auto blob = g.to_bytes();
        auto e1 = g.timesteps(20, 1.0);
        auto g2 = qmcb::DefaultQmcIsingGraph::from_bytes(blob);
        auto e2 = g2.timesteps(20, 1.0);
        ASSERT(e1 == e2 && g.state_ref() == g2.state_ref() && g.rng_cursors() == g2.rng_cursors());
        // imaginary_time_fold with a closure: the propagated state returns to the p = 0 state only at the end, and
        // the number of slices seen equals the cutoff
        size_t slices = g2.imaginary_time_fold(0, [](size_t acc, const std::vector<bool> &) { return acc + 1; }, (size_t)0);
        ASSERT(slices == g2.get_cutoff()[0]);
        auto ac = g2.calculate_variable_autocorrelation(64, 1.0, 1);
        ASSERT(ac.size() == 3 && ac[0].size() == 64 && std::abs(ac[0][0] - 1.0) < 1e-12);
    }
    {  // tests/check_loop_crash.rs:6-74 run_single_bond / run_double_bond: 100 loop updates on a hand-built string, verify
        std::vector<double> swap(16, 0.0);  // weight 1 when inputs == outputs or inputs == reversed outputs (:19-27)
        for (int o0 = 0; o0 < 2; o0++)
            for (int o1 = 0; o1 < 2; o1++)
                for (int i0 = 0; i0 < 2; i0++)
                    for (int i1 = 0; i1 < 2; i1++)
                        if ((i0 == o0 && i1 == o1) || (i0 == o1 && i1 == o0)) swap[(o0 << 3) | (o1 << 2) | (i0 << 1) | i1] = 1.0;
        for (size_t nbonds : {(size_t)1, (size_t)2}) {
            std::vector<bool> st(nbonds + 1, false);
            qmcb::Qmc q(nbonds + 1, seeds, true, &st);
            for (size_t b = 0; b < nbonds; b++) q.make_interaction(swap, {b, b + 1});
            q.increase_cutoff_to(nbonds);
            q.build();
            std::vector<uint32_t> ops;
            for (size_t b = 0; b < nbonds; b++) ops.push_back((uint32_t)b);  // FastOp::diagonal([b, b+1], b, [false, false])
            for (size_t r = 0; r < seeds.size(); r++) q.load_ops(r, ops, &st);
            for (int i = 0; i < 100; i++) q.loop_update();
            ASSERT(q.verify());
            bool moved = false;
            for (auto &s : q.state_ref())
                for (bool b : s) moved = moved || b;
            ASSERT(moved);
        }
        // a model that moves by loop updates only (exchange terms, no cluster edges): Qmc::timestep with do_loop_updates
        auto xxz = [](double d0, double d1, double x) {
            std::vector<double> m(16, 0.0);
            m[0] = m[15] = d0, m[5] = m[10] = d1, m[6] = m[9] = x;
            return m;
        };
        qmcb::Qmc q(4, seeds, true);
        q.make_interaction(xxz(0.4, 1.1, 0.8), {0, 1}), q.make_interaction(xxz(0.9, 0.5, 0.6), {1, 2});
        q.make_interaction(xxz(0.3, 1.0, 1.0), {2, 3}), q.make_interaction(xxz(0.7, 0.7, 0.5), {3, 0});
        auto e = q.timesteps(500, 1.2);
        ASSERT(q.verify() && e.size() == seeds.size());
        for (double x : e) ASSERT(x < 0.0);
    }
    {  // error behaviour: status codes become exceptions, never aborts
        bool threw = false;
        try {
            qmcb::DefaultQmcIsingGraph::new_with_rng({{{0, 0}, 1.0}}, 1.0, 0.0, 2, {1});  // self-loop
        } catch (const qmcb::Error &e) {
            threw = e.code == QMCB_ERR_BAD_ARG;
        }
        ASSERT(threw);
    }
    {  // classical mirror
        std::vector<Edge> edges;
        const size_t L = 8;
        for (size_t i = 0; i < L; i++)
            for (size_t j = 0; j < L; j++) {
                edges.push_back({{j * L + i, j * L + (i + 1) % L}, -1.0});
                edges.push_back({{j * L + i, ((j + 1) % L) * L + i}, -1.0});
            }
        auto g = qmcb::GraphState::create(edges, std::vector<double>(L * L, 0.0), {1, 2, 3}, {2.0, 2.0, 2.0});
        g.sweeps(200);
        for (double e : g.get_energy()) ASSERT(e < -80.0);  // quench into the ordered phase: -128 (uniform) or -96 (one stripe)
        for (int t = 0; t < 5; t++) g.do_time_step();       // the reference's own schedule keeps a cold system cold
        for (double e : g.get_energy()) ASSERT(e < -80.0);
    }
    {  // classical/graph.rs:481-498 test_worm_flip and :564-581 test_worm_flip_doubles, on every replica
        std::vector<Edge> tri = {{{0, 1}, 1.0}, {{1, 2}, 1.0}, {{2, 0}, 1.0}};
        std::vector<bool> zeros(3, false);
        std::vector<uint64_t> keys = {1, 2, 3, 4, 5, 6, 7, 8};
        std::vector<double> betas(keys.size(), 1.0);
        auto g = qmcb::GraphState::create(tri, {0., 0., 0.}, keys, betas, 0, &zeros);
        g.do_worm_flip(1, false);
        for (auto &s : g.state_ref()) ASSERT(s[0] && s[1] && s[2]);
        auto g2 = qmcb::GraphState::create(tri, {0., 0., 0.}, keys, betas, 0, &zeros);
        g2.do_worm_flip(1, true);
        for (auto &s : g2.state_ref()) ASSERT(s[0] == s[1] && s[1] == s[2]);
    }
    std::printf("cpp mirror ok\n");
    return 0;
}
