// qmcb.hpp -- header-only C++ host side above the C ABI (include/qmcb.h), mirroring the reference's
// interface for the hot path with the same names and argument meaning:
//   qmc::sse::QmcIsingGraph + QmcStepper   (src/sse/qmc_ising.rs:131-148, qmc_traits/qmc_stepper.rs:2-168)
//   qmc::classical::graph::GraphState      (src/classical/graph.rs:56-88, 350-447)
// One object is a BATCH of replicas (one rng key and one beta per replica); wherever the reference
// returns one value per graph these return one per replica.  Reference panics / Result<(), String>
// become qmcb::Error exceptions carrying qmcb_last_error().  The reference is Rust; this image has no
// Rust toolchain, so this C++ mirror (and the Python one) is what the parity tests are written against.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "qmcb.h"

namespace qmcb {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc != QMCB_OK) throw Error(rc, qmcb_last_error());
}

using Edge = std::pair<std::pair<size_t, size_t>, double>;  // ((vara, varb), J) as in the reference

namespace detail {
struct Lat {
    std::vector<uint32_t> va, vb;
    std::vector<double> j;
    QmcbLattice l{};
    Lat(const std::vector<Edge> &edges, double transverse, double longitudinal, size_t nvars_hint = 0) {
        size_t nv = nvars_hint;
        for (auto &e : edges) {
            va.push_back((uint32_t)e.first.first), vb.push_back((uint32_t)e.first.second), j.push_back(e.second);
            nv = std::max(nv, std::max(e.first.first, e.first.second) + 1);  // qmc_ising.rs:92
        }
        l.nvars = (uint32_t)nv, l.nedges = (uint32_t)edges.size();
        l.va = va.data(), l.vb = vb.data(), l.J = j.data();
        l.transverse = transverse, l.longitudinal = longitudinal;
    }
};
}  // namespace detail

class QmcIsingGraph {
    QmcbHandle *h_ = nullptr;
    size_t nvars_ = 0, replicas_ = 0;
    std::vector<double> betas_;

    void set_beta(double beta) {
        if (betas_.empty() || betas_[0] != beta || betas_.size() != replicas_) {
            betas_.assign(replicas_, beta);
            check(qmcb_set_betas(h_, betas_.data()));
        }
    }

public:
    // QmcIsingGraph::new_with_rng (qmc_ising.rs:131-148): `rng` becomes one Philox key per replica;
    // state = nullptr draws the spins from each stream as make_random_spin_state does.
    static QmcIsingGraph new_with_rng(const std::vector<Edge> &edges, double transverse, double longitudinal, size_t cutoff,
                                      const std::vector<uint64_t> &rng_keys, const std::vector<bool> *state = nullptr,
                                      int mode = QMCB_MODE_STRICT, int device = 0) {
        QmcIsingGraph g;
        detail::Lat lat(edges, transverse, longitudinal);
        g.nvars_ = lat.l.nvars, g.replicas_ = rng_keys.size();
        std::vector<double> betas(rng_keys.size(), 1.0);
        std::vector<uint8_t> init;
        if (state)
            for (size_t r = 0; r < rng_keys.size(); r++)
                for (bool b : *state) init.push_back(b ? 1 : 0);
        check(qmcb_create(&lat.l, (uint32_t)rng_keys.size(), betas.data(), rng_keys.data(), cutoff, 0, state ? init.data() : nullptr,
                          device, &g.h_));
        check(qmcb_set_mode(g.h_, mode));
        return g;
    }
    QmcIsingGraph() = default;
    QmcIsingGraph(QmcIsingGraph &&o) noexcept { *this = std::move(o); }
    QmcIsingGraph &operator=(QmcIsingGraph &&o) noexcept {
        std::swap(h_, o.h_), std::swap(nvars_, o.nvars_), std::swap(replicas_, o.replicas_), std::swap(betas_, o.betas_);
        return *this;
    }
    QmcIsingGraph(const QmcIsingGraph &) = delete;
    ~QmcIsingGraph() { qmcb_destroy(h_); }

    // ---- QmcStepper -----------------------------------------------------------------------
    std::vector<std::vector<bool>> timestep(double beta) {  // qmc_stepper.rs:4
        set_beta(beta);
        check(qmcb_timesteps(h_, 1, 1, nullptr, nullptr));
        return state_ref();
    }
    std::vector<double> timesteps(size_t t, double beta) {  // :17-20
        set_beta(beta);
        std::vector<double> e(replicas_);
        check(qmcb_timesteps(h_, t, 1, e.data(), nullptr));
        return e;
    }
    // :23-40; samples[replica][k][var]
    std::pair<std::vector<std::vector<std::vector<bool>>>, std::vector<double>> timesteps_sample(size_t t, double beta, size_t sampling_freq = 1) {
        set_beta(beta);
        const size_t k = t / sampling_freq;
        std::vector<double> e(replicas_);
        std::vector<uint8_t> raw(replicas_ * k * nvars_ + 1);
        check(qmcb_timesteps(h_, t, sampling_freq, e.data(), raw.data()));
        std::vector<std::vector<std::vector<bool>>> s(replicas_, std::vector<std::vector<bool>>(k, std::vector<bool>(nvars_)));
        for (size_t r = 0; r < replicas_; r++)
            for (size_t i = 0; i < k; i++)
                for (size_t v = 0; v < nvars_; v++) s[r][i][v] = raw[(r * k + i) * nvars_ + v] != 0;
        return {s, e};
    }
    void single_diagonal_step(double beta) {  // qmc_ising.rs:208-270
        set_beta(beta);
        check(qmcb_single_diagonal_step(h_));
    }
    std::vector<uint64_t> single_cluster_step() {  // qmc_ising.rs:273-320
        std::vector<uint64_t> n(replicas_);
        check(qmcb_single_cluster_step(h_, n.data()));
        return n;
    }
    std::vector<uint64_t> get_n() {
        std::vector<uint64_t> n(replicas_);
        check(qmcb_get_n(h_, n.data()));
        return n;
    }
    std::vector<std::vector<bool>> state_ref() {
        std::vector<uint8_t> raw(replicas_ * nvars_);
        check(qmcb_get_states(h_, raw.data()));
        std::vector<std::vector<bool>> s(replicas_, std::vector<bool>(nvars_));
        for (size_t r = 0; r < replicas_; r++)
            for (size_t v = 0; v < nvars_; v++) s[r][v] = raw[r * nvars_ + v] != 0;
        return s;
    }
    double get_energy_for_average_n(double average_n, double beta) const { return -(average_n / beta) + get_offset(); }  // :805-809
    // ---- accessors (qmc_ising.rs:496-560) ---------------------------------------------------
    std::vector<uint64_t> get_cutoff() {
        std::vector<uint64_t> c(replicas_);
        check(qmcb_get_cutoffs(h_, c.data()));
        return c;
    }
    void set_cutoff(size_t cutoff) {
        for (size_t r = 0; r < replicas_; r++) check(qmcb_set_cutoff(h_, (uint32_t)r, cutoff));
    }
    size_t get_nvars() const { return nvars_; }
    size_t num_replicas() const { return replicas_; }
    double get_offset() const {
        double o = 0;
        check(qmcb_get_offset(h_, &o));
        return o;
    }
    std::vector<uint64_t> get_bond_counts(size_t r) {
        uint32_t nb = 0;
        check(qmcb_num_bonds(h_, &nb));
        std::vector<uint64_t> c(nb);
        check(qmcb_get_bond_counts(h_, (uint32_t)r, c.data()));
        return c;
    }
    bool verify() {  // Verify::verify, qmc_ising.rs:829-860
        for (size_t r = 0; r < replicas_; r++) {
            int ok = 0;
            check(qmcb_verify(h_, (uint32_t)r, &ok));
            if (!ok) return false;
        }
        return true;
    }
    void set_mode(int mode) { check(qmcb_set_mode(h_, mode)); }
    void set_enable_heatbath(bool enable) { check(qmcb_set_enable_heatbath(h_, enable ? 1 : 0)); }  // qmc_ising.rs:444-486
    QmcbHandle *raw() { return h_; }
};
using DefaultQmcIsingGraph = QmcIsingGraph;

class GraphState {
    CmcbHandle *h_ = nullptr;
    size_t nvars_ = 0, replicas_ = 0;

public:
    // GraphState::new (graph.rs:56-59); betas per replica because the sweep needs them at creation
    static GraphState create(const std::vector<Edge> &edges, const std::vector<double> &biases, const std::vector<uint64_t> &rng_keys,
                             const std::vector<double> &betas, int device = 0) {
        GraphState g;
        detail::Lat lat(edges, 0.0, 0.0, biases.size());
        g.nvars_ = biases.size(), g.replicas_ = rng_keys.size();
        check(cmcb_create(&lat.l, biases.data(), (uint32_t)rng_keys.size(), betas.data(), rng_keys.data(), nullptr, device, &g.h_));
        return g;
    }
    GraphState() = default;
    GraphState(GraphState &&o) noexcept { std::swap(h_, o.h_), std::swap(nvars_, o.nvars_), std::swap(replicas_, o.replicas_); }
    GraphState(const GraphState &) = delete;
    ~GraphState() { cmcb_destroy(h_); }
    void do_time_step(size_t nsweeps = 1) { check(cmcb_sweeps(h_, nsweeps)); }  // graph.rs:350-406 (checkerboard schedule)
    std::vector<double> get_energy() {                                           // graph.rs:430-447
        std::vector<double> e(replicas_);
        check(cmcb_energy(h_, e.data()));
        return e;
    }
    std::vector<std::vector<bool>> state_ref() {
        std::vector<uint8_t> raw(replicas_ * nvars_);
        check(cmcb_get_states(h_, raw.data()));
        std::vector<std::vector<bool>> s(replicas_, std::vector<bool>(nvars_));
        for (size_t r = 0; r < replicas_; r++)
            for (size_t v = 0; v < nvars_; v++) s[r][v] = raw[r * nvars_ + v] != 0;
        return s;
    }
};

}  // namespace qmcb
