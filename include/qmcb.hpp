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
    // get_transverse_field / get_longitudinal_field / clone_state (qmc_ising.rs:502-529)
    double get_transverse_field() const {
        double t = 0, l = 0;
        check(qmcb_get_fields(h_, &t, &l));
        return t;
    }
    double get_longitudinal_field() const {
        double t = 0, l = 0;
        check(qmcb_get_fields(h_, &t, &l));
        return l;
    }
    std::vector<std::vector<bool>> clone_state() { return state_ref(); }
    void set_mode(int mode) { check(qmcb_set_mode(h_, mode)); }
    void set_enable_heatbath(bool enable) { check(qmcb_set_enable_heatbath(h_, enable ? 1 : 0)); }  // qmc_ising.rs:444-486
    // RVB update (rvb.rs:60-291): set_run_rvb qmc_ising.rs:434-441, single_rvb_sweep :322-420 (successes per replica,
    // attempts), rvb_success_rate :604-607
    void set_run_rvb(bool run_rvb) { check(qmcb_set_run_rvb(h_, run_rvb ? 1 : 0)); }
    std::pair<std::vector<uint64_t>, uint64_t> single_rvb_sweep(int64_t updates_in_sweep = -1) {
        std::vector<uint64_t> succ(replicas_);
        uint64_t attempts = 0;
        check(qmcb_single_rvb_sweep(h_, updates_in_sweep, succ.data(), &attempts));
        return {succ, attempts};
    }
    std::vector<double> rvb_success_rate() {
        std::vector<double> rate(replicas_);
        check(qmcb_rvb_success_rate(h_, rate.data(), nullptr, nullptr));
        return rate;
    }
    // graphs with their own couplings in one batch (tempering_traits.rs:122-154): rows of (J[E], transverse, longitudinal)
    void set_hamiltonians(const std::vector<std::vector<double>> &j_rows, const std::vector<double> &transverse,
                          const std::vector<double> &longitudinal, const std::vector<uint32_t> &ham_of_replica) {
        std::vector<double> flat;
        for (auto &row : j_rows) flat.insert(flat.end(), row.begin(), row.end());
        check(qmcb_set_hamiltonians(h_, (uint32_t)transverse.size(), flat.data(), transverse.data(), longitudinal.data(), ham_of_replica.data()));
    }
    // QmcStepper::imaginary_time_fold (qmc_stepper.rs:165-168) for replica r, the closure evaluated on the host
    template <typename F, typename T>
    T imaginary_time_fold(size_t r, F fold_fn, T init) {
        std::vector<uint64_t> cut(replicas_);
        check(qmcb_get_cutoffs(h_, cut.data()));
        std::vector<uint8_t> raw(nvars_);
        std::vector<bool> st(nvars_);
        T acc = init;
        for (uint64_t p = 0; p < cut[r]; p++) {
            check(qmcb_itime_state(h_, (uint32_t)r, p, raw.data()));
            for (size_t v = 0; v < nvars_; v++) st[v] = raw[v] != 0;
            acc = fold_fn(acc, st);
        }
        return acc;
    }
    // QmcAutoCorrelations::calculate_variable_autocorrelation (autocorrelations.rs:48-61): [replica][lag]
    std::vector<std::vector<double>> calculate_variable_autocorrelation(size_t timesteps, double beta, size_t sampling_freq = 1) {
        set_beta(beta);
        const size_t T = timesteps / sampling_freq;
        std::vector<double> flat(replicas_ * T);
        check(qmcb_variable_autocorrelation(h_, timesteps, sampling_freq, flat.data(), nullptr, nullptr));
        std::vector<std::vector<double>> out(replicas_);
        for (size_t r = 0; r < replicas_; r++) out[r].assign(flat.begin() + r * T, flat.begin() + (r + 1) * T);
        return out;
    }
    // serde replacement (qmc_ising.rs:1001-1087): the whole batch, stream positions included
    std::vector<uint8_t> to_bytes() {
        uint64_t n = 0;
        check(qmcb_checkpoint_size(h_, &n));
        std::vector<uint8_t> buf(n);
        check(qmcb_checkpoint_save(h_, buf.data(), n));
        return buf;
    }
    static QmcIsingGraph from_bytes(const std::vector<uint8_t> &buf, int device = 0) {
        QmcIsingGraph g;
        check(qmcb_checkpoint_load(buf.data(), buf.size(), device, &g.h_));
        uint32_t r = 0, n = 0;
        check(qmcb_num_replicas(g.h_, &r));
        check(qmcb_num_vars(g.h_, &n));
        g.replicas_ = r, g.nvars_ = n;
        g.betas_.resize(r);
        check(qmcb_get_betas(g.h_, g.betas_.data()));
        return g;
    }
    std::vector<uint64_t> rng_cursors() {
        std::vector<uint64_t> c(replicas_);
        check(qmcb_get_rng_cursors(h_, c.data()));
        return c;
    }
    QmcbHandle *raw() { return h_; }
    // FastOps::new_from_ops (fast_ops.rs:80-87) for replica r: op words as include/qmcb.h packs them
    void load_ops(size_t r, const std::vector<uint32_t> &words, const std::vector<bool> *state = nullptr) {
        std::vector<uint8_t> st;
        if (state)
            for (bool b : *state) st.push_back(b ? 1 : 0);
        check(qmcb_load_ops(h_, (uint32_t)r, words.data(), words.size(), state ? st.data() : nullptr));
    }
    std::vector<uint32_t> dump_ops(size_t r) {
        std::vector<uint64_t> cut(replicas_);
        check(qmcb_get_cutoffs(h_, cut.data()));
        std::vector<uint32_t> w(cut[r]);
        check(qmcb_dump_ops(h_, (uint32_t)r, w.data(), w.size()));
        return w;
    }

protected:
    void adopt(QmcbHandle *h, size_t nvars, size_t replicas) { h_ = h, nvars_ = nvars, replicas_ = replicas; }
    bool has_handle() const { return h_ != nullptr; }
};
using DefaultQmcIsingGraph = QmcIsingGraph;

// qmc::sse::Qmc (qmc_runner.rs:22-403): interactions given as matrices (Interaction::at indexing, :560-664); the batch is
// built when the first step needs it.  timestep = diagonal update, loop update when do_loop_updates (directed_loop.rs:
// 103-301), cluster update with Ising symmetry, free bits (:363-377).  Every QmcIsingGraph accessor works on it.
class Qmc : public QmcIsingGraph {
    struct Bond {
        std::vector<double> mat;
        std::vector<uint32_t> vars;
    };
    std::vector<Bond> bonds_;
    std::vector<uint64_t> keys_;
    std::vector<uint8_t> init_;
    size_t nv_ = 0, cutoff_ = 0;
    double offset_ = 0.0;
    bool loops_ = false;
    int mode_ = QMCB_MODE_STRICT, device_ = 0;

    void add(std::vector<double> mat, const std::vector<size_t> &vars, bool diagonal, bool and_offset) {
        if (has_handle()) throw Error(QMCB_ERR_UNSUPPORTED, "interactions are fixed once the batch has been stepped");
        const size_t n = vars.size(), tn = (size_t)1 << n;
        if (mat.size() != (diagonal ? tn : tn * tn)) throw Error(QMCB_ERR_BAD_ARG, "Given vars do not match the matrix size");
        if (and_offset) {  // new_offset :508-521 / new_diagonal_offset :424-436
            const size_t stride = diagonal ? 1 : tn + 1;
            double lo = mat[0];
            for (size_t k = 0; k < tn; k++) lo = std::min(lo, mat[k * stride]);
            for (size_t k = 0; k < tn; k++) mat[k * stride] -= lo;
            offset_ -= lo;
        }
        Bond b;
        b.mat = std::move(mat);
        for (size_t v : vars) b.vars.push_back((uint32_t)v);
        bonds_.push_back(std::move(b));
    }

public:
    // Qmc::new / new_with_state (qmc_runner.rs:48-87): cutoff = nvars, one Philox key per replica
    Qmc(size_t nvars, const std::vector<uint64_t> &rng_keys, bool do_loop_updates, const std::vector<bool> *state = nullptr,
        int mode = QMCB_MODE_STRICT, int device = 0)
        : keys_(rng_keys), nv_(nvars), cutoff_(nvars), loops_(do_loop_updates), mode_(mode), device_(device) {
        if (state)
            for (size_t r = 0; r < rng_keys.size(); r++)
                for (bool b : *state) init_.push_back(b ? 1 : 0);
    }
    void make_interaction(std::vector<double> mat, const std::vector<size_t> &vars) { add(std::move(mat), vars, false, false); }             // :113-122
    void make_interaction_and_offset(std::vector<double> mat, const std::vector<size_t> &vars) { add(std::move(mat), vars, false, true); }   // :125-135
    void make_diagonal_interaction(std::vector<double> mat, const std::vector<size_t> &vars) { add(std::move(mat), vars, true, false); }     // :138-146
    void make_diagonal_interaction_and_offset(std::vector<double> mat, const std::vector<size_t> &vars) { add(std::move(mat), vars, true, true); }  // :149-156
    void increase_cutoff_to(size_t cutoff) { cutoff_ = std::max(cutoff_, cutoff); }  // :307-309 (before the first step)
    // builds the batch; called by the first step
    void build() {
        if (has_handle()) return;
        std::vector<uint32_t> nv, vars, len;
        std::vector<double> mats;
        for (auto &b : bonds_) {
            nv.push_back((uint32_t)b.vars.size()), len.push_back((uint32_t)b.mat.size());
            vars.push_back(b.vars[0]), vars.push_back(b.vars.size() > 1 ? b.vars[1] : 0);
            mats.insert(mats.end(), b.mat.begin(), b.mat.end());
        }
        QmcbInteractions in{(uint32_t)nv_, (uint32_t)bonds_.size(), nv.data(), vars.data(), len.data(), mats.data(), offset_, loops_ ? 1 : 0};
        std::vector<double> betas(keys_.size(), 1.0);
        QmcbHandle *h = nullptr;
        check(qmcb_create_qmc(&in, (uint32_t)keys_.size(), betas.data(), keys_.data(), cutoff_, 0, init_.empty() ? nullptr : init_.data(), device_, &h));
        adopt(h, nv_, keys_.size());
        set_mode(mode_);
    }
    void loop_update() {  // Qmc::loop_update :205-220
        build();
        check(qmcb_loop_update(raw()));
    }
    void set_do_loop_updates(bool enable) {  // :268-270
        loops_ = enable;
        if (has_handle()) check(qmcb_set_do_loop_updates(raw(), enable ? 1 : 0));
    }
    bool should_do_loop_update() const { return loops_; }
    std::vector<double> timesteps(size_t t, double beta) {
        build();
        return QmcIsingGraph::timesteps(t, beta);
    }
    std::vector<std::vector<bool>> timestep(double beta) {
        build();
        return QmcIsingGraph::timestep(beta);
    }
};

// qmc::sse::parallel_tempering::TemperingContainer (tempering_container.rs:19-302) on one GPU: n_chains ladders of
// betas.size() slots; add_qmc_stepper is replaced by giving the whole ladder at construction.
class TemperingContainer {
    QmcIsingGraph g_;
    size_t n_betas_ = 0, n_chains_ = 0;

public:
    TemperingContainer(const std::vector<Edge> &edges, double transverse, double longitudinal, size_t cutoff, const std::vector<double> &betas,
                       size_t n_chains, const std::vector<uint64_t> &rng_keys, uint64_t pt_key, int mode = QMCB_MODE_FAST)
        : g_(QmcIsingGraph::new_with_rng(edges, transverse, longitudinal, cutoff, rng_keys, nullptr, mode)), n_betas_(betas.size()), n_chains_(n_chains) {
        std::vector<double> all;
        for (size_t c = 0; c < n_chains; c++) all.insert(all.end(), betas.begin(), betas.end());
        if (all.size() != rng_keys.size()) throw Error(QMCB_ERR_BAD_ARG, "one rng key per slot");
        check(qmcb_pt_configure(g_.raw(), (uint32_t)n_chains, (uint32_t)betas.size(), 0, all.data(), rng_keys.data(), pt_key));
    }
    size_t num_graphs() const { return n_betas_ * n_chains_; }
    void timesteps(size_t t) { check(qmcb_timesteps(g_.raw(), t, 1, nullptr, nullptr)); }  // :76-81
    void tempering_step() { check(qmcb_pt_step(g_.raw())); }                                // :121-149 (:373-402)
    // timesteps_sample / parallel_timesteps_sample (:166-208, :411-453): (states, energy_acc) per slot
    std::vector<std::pair<std::vector<std::vector<bool>>, double>> timesteps_sample(size_t timesteps, size_t replica_swap_freq, size_t sampling_freq) {
        const size_t S = num_graphs(), T = timesteps / sampling_freq, N = g_.get_nvars();
        std::vector<double> energy(S);
        std::vector<uint8_t> raw(S * T * N);
        std::vector<uint32_t> slots(S * T);
        check(qmcb_pt_timesteps_sample(g_.raw(), timesteps, replica_swap_freq, sampling_freq, energy.data(), raw.data(), slots.data()));
        std::vector<std::pair<std::vector<std::vector<bool>>, double>> out(S);
        for (size_t s = 0; s < S; s++) out[s].second = energy[s];
        for (size_t k = 0; k < T; k++)
            for (size_t c = 0; c < S; c++) {
                std::vector<bool> st(N);
                for (size_t v = 0; v < N; v++) st[v] = raw[(c * T + k) * N + v] != 0;
                out[slots[c * T + k]].first.push_back(std::move(st));
            }
        return out;
    }
    uint64_t get_total_swaps() {                                                            // :231-233
        uint64_t s = 0;
        check(qmcb_pt_total_swaps(g_.raw(), &s));
        return s;
    }
    // ParallelTemperingAutocorrelations / ParallelTemperingBondAutoCorrelations (:484-630): one autocorrelation per slot
    std::vector<std::vector<double>> calculate_variable_autocorrelation(size_t timesteps, size_t replica_swap_freq = 1, size_t sampling_freq = 1) {
        const size_t S = n_betas_ * n_chains_, T = timesteps / sampling_freq;
        std::vector<double> ac(S * T);
        check(qmcb_pt_variable_autocorrelation(g_.raw(), timesteps, replica_swap_freq, sampling_freq, ac.data(), nullptr, nullptr));
        return split(ac, S, T);
    }
    std::vector<std::vector<double>> calculate_spin_product_autocorrelation(size_t timesteps, size_t replica_swap_freq,
                                                                            const std::vector<std::vector<uint32_t>> &var_products, size_t sampling_freq = 1) {
        const size_t S = n_betas_ * n_chains_, T = timesteps / sampling_freq;
        std::vector<uint32_t> off(1, 0), flat;
        for (auto &p : var_products) flat.insert(flat.end(), p.begin(), p.end()), off.push_back((uint32_t)flat.size());
        std::vector<double> ac(S * T);
        check(qmcb_pt_spin_product_autocorrelation(g_.raw(), timesteps, replica_swap_freq, sampling_freq, (uint32_t)var_products.size(), off.data(),
                                                   flat.data(), ac.data(), nullptr, nullptr));
        return split(ac, S, T);
    }
    std::vector<std::vector<double>> calculate_bond_autocorrelation(size_t timesteps, size_t replica_swap_freq = 1, size_t sampling_freq = 1) {
        const size_t S = n_betas_ * n_chains_, T = timesteps / sampling_freq;
        std::vector<double> ac(S * T);
        check(qmcb_pt_bond_autocorrelation(g_.raw(), timesteps, replica_swap_freq, sampling_freq, ac.data(), nullptr, nullptr));
        return split(ac, S, T);
    }
    bool verify() { return g_.verify(); }
    QmcIsingGraph &graph() { return g_; }

private:
    static std::vector<std::vector<double>> split(const std::vector<double> &flat, size_t S, size_t T) {
        std::vector<std::vector<double>> out(S);
        for (size_t s = 0; s < S; s++) out[s].assign(flat.begin() + s * T, flat.begin() + (s + 1) * T);
        return out;
    }
};

class GraphState {
    CmcbHandle *h_ = nullptr;
    size_t nvars_ = 0, replicas_ = 0;

public:
    // GraphState::new (graph.rs:56-59); betas per replica because the sweep needs them at creation
    static GraphState create(const std::vector<Edge> &edges, const std::vector<double> &biases, const std::vector<uint64_t> &rng_keys,
                             const std::vector<double> &betas, int device = 0, const std::vector<bool> *state = nullptr) {
        GraphState g;
        detail::Lat lat(edges, 0.0, 0.0, biases.size());
        g.nvars_ = biases.size(), g.replicas_ = rng_keys.size();
        std::vector<uint8_t> init;  // new_with_state_and_rng (graph.rs:62-88): the same state for every replica
        if (state)
            for (size_t r = 0; r < g.replicas_; r++)
                for (bool b : *state) init.push_back(b);
        check(cmcb_create(&lat.l, biases.data(), (uint32_t)rng_keys.size(), betas.data(), rng_keys.data(), state ? init.data() : nullptr,
                          device, &g.h_));
        return g;
    }
    GraphState() = default;
    GraphState(GraphState &&o) noexcept { std::swap(h_, o.h_), std::swap(nvars_, o.nvars_), std::swap(replicas_, o.replicas_); }
    GraphState(const GraphState &) = delete;
    ~GraphState() { cmcb_destroy(h_); }
    void sweeps(size_t nsweeps = 1) { check(cmcb_sweeps(h_, nsweeps)); }  // checkerboard schedule, the throughput path
    // GraphState::do_time_step (graph.rs:350-406), the reference's own schedule; SIZE_MAX = None.  Returns the move drawn per replica.
    std::vector<uint8_t> do_time_step(size_t nspinupdates = SIZE_MAX, size_t nedgeupdates = SIZE_MAX, size_t nwormupdates = SIZE_MAX,
                                      bool only_basic_moves = false) {
        std::vector<uint8_t> choice(replicas_);
        check(cmcb_do_time_step(h_, nspinupdates, nedgeupdates, nwormupdates, only_basic_moves, choice.data()));
        return choice;
    }
    void do_worm_flip(size_t count, bool allow_doubles) { check(cmcb_worm_flips(h_, count, allow_doubles)); }  // graph.rs:179-318
    void enable_edge_importance_sampling(bool enable) { check(cmcb_enable_edge_importance_sampling(h_, enable)); }  // :321-336
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
