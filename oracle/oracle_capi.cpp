// ORACLE (test infrastructure, not product): C entry points over oracle/cpu_backend.hpp so the
// Python tests can diff libcm31 against the CPU restatement op by op.  Same flattened argument
// conventions as include/cm31.h, but every pointer is a HOST pointer.
#include <cstdio>

#include "cpu_backend.hpp"

using namespace orc;

static std::vector<M31> to_m31(const u32* p, size_t n) {
    std::vector<M31> v(n);
    for (size_t i = 0; i < n; i++) v[i] = M31((u64)p[i]);
    return v;
}
static QM31 qm(const u32* p) { return QM31::from_u32(p[0], p[1], p[2], p[3]); }
static QPoint qpt(const u32* p) { return QPoint{qm(p), qm(p + 4)}; }

extern "C" {

u32 orc_m31_add(u32 a, u32 b) { return (M31((u64)a) + M31((u64)b)).v; }
u32 orc_m31_sub(u32 a, u32 b) { return (M31((u64)a) - M31((u64)b)).v; }
u32 orc_m31_mul(u32 a, u32 b) { return (M31((u64)a) * M31((u64)b)).v; }
u32 orc_m31_inv(u32 a) { return M31((u64)a).inverse().v; }
void orc_qm31_mul(const u32* a, const u32* b, u32* out) { (qm(a) * qm(b)).to_u32(out); }
void orc_qm31_inv(const u32* a, u32* out) { qm(a).inverse().to_u32(out); }
void orc_point_from_index(u32 index, u32* out) {
    Point p = point_from_index(index);
    out[0] = p.x.v;
    out[1] = p.y.v;
}
void orc_domain_at(u32 log_size, u32 i, u32* out) {
    Point p = ODomain::canonic(log_size).at(i);
    out[0] = p.x.v;
    out[1] = p.y.v;
}

void orc_blake2s(const uint8_t* data, size_t len, uint8_t* out) {
    OBlake2s h;
    h.update(data, len);
    h.finalize(out);
}

void orc_twiddles(u32 log_size, u32* tw_out, u32* itw_out) {
    OTwiddles t = precompute_twiddles(log_size);
    for (size_t i = 0; i < t.tw.size(); i++) {
        tw_out[i] = t.tw[i].v;
        itw_out[i] = t.itw[i].v;
    }
}

// in place: values (2^log_size) -> coefficients
void orc_interpolate(u32* values, u32 log_size, u32 n_cols) {
    OTwiddles t = precompute_twiddles(log_size < 3 ? 3 : log_size);
    size_t n = (size_t)1 << log_size;
#pragma omp parallel for schedule(dynamic)
    for (u32 c = 0; c < n_cols; c++) {
        std::vector<M31> v = to_m31(values + c * n, n);
        interpolate(v, t);
        for (size_t i = 0; i < n; i++) values[c * n + i] = v[i].v;
    }
}
void orc_evaluate(const u32* coeffs, u32 log_size, u32 log_eval, u32 n_cols, u32* out) {
    OTwiddles t = precompute_twiddles(log_eval < 3 ? 3 : log_eval);
    size_t n = (size_t)1 << log_size, m = (size_t)1 << log_eval;
#pragma omp parallel for schedule(dynamic)
    for (u32 c = 0; c < n_cols; c++) {
        std::vector<M31> v = evaluate(to_m31(coeffs + c * n, n), log_eval, t);
        for (size_t i = 0; i < m; i++) out[c * m + i] = v[i].v;
    }
}
void orc_eval_at_point(const u32* coeffs, u32 log_size, const u32* point, u32* out) {
    eval_at_point(to_m31(coeffs, (size_t)1 << log_size), qpt(point)).to_u32(out);
}

// cols: n_cols contiguous columns of 2^log_size; prev: 2^(log_size+1) hashes or NULL
void orc_commit_on_layer(u32 log_size, const uint8_t* prev, const u32* cols, u32 n_cols, uint8_t* out) {
    size_t n = (size_t)1 << log_size;
    std::vector<std::vector<M31>> cv(n_cols);
    std::vector<const std::vector<M31>*> cp;
    for (u32 c = 0; c < n_cols; c++) {
        cv[c] = to_m31(cols + c * n, n);
        cp.push_back(&cv[c]);
    }
    std::vector<Hash> pv;
    if (prev) {
        pv.resize(2 * n);
        memcpy(pv.data(), prev, 64 * n);
    }
    std::vector<Hash> o = commit_on_layer(log_size, prev ? &pv : nullptr, cp);
    memcpy(out, o.data(), 32 * n);
}

static SecureColumn sc_from(const u32* src4, size_t n) {  // 4 contiguous coordinate columns
    SecureColumn s(n);
    for (int k = 0; k < 4; k++)
        for (size_t i = 0; i < n; i++) s.c[k][i] = M31((u64)src4[k * n + i]);
    return s;
}
static void sc_to(const SecureColumn& s, u32* dst4) {
    size_t n = s.size();
    for (int k = 0; k < 4; k++)
        for (size_t i = 0; i < n; i++) dst4[k * n + i] = s.c[k][i].v;
}
void orc_fold_line(const u32* src4, u32 log_size, const u32* alpha, u32* dst4) {
    sc_to(fold_line(sc_from(src4, (size_t)1 << log_size), qm(alpha)), dst4);
}
void orc_fold_circle_into_line(u32* dst4, const u32* src4, u32 log_size, const u32* alpha) {
    size_t n = (size_t)1 << log_size;
    SecureColumn d = sc_from(dst4, n / 2);
    fold_circle_into_line(d, sc_from(src4, n), qm(alpha));
    sc_to(d, dst4);
}
void orc_decompose(const u32* src4, u32 log_size, u32* dst4, u32* lambda_out) {
    QM31 lam;
    sc_to(decompose(sc_from(src4, (size_t)1 << log_size), &lam), dst4);
    lam.to_u32(lambda_out);
}
void orc_accumulate_quotients(u32 log_size, const u32* cols, u32 n_cols, const u32* random_coeff, u32 n_batches,
                              const u32* batch_points, const u32* batch_start, const u32* col_idx, const u32* values,
                              u32* out4) {
    size_t n = (size_t)1 << log_size;
    std::vector<std::vector<M31>> cv(n_cols);
    std::vector<const std::vector<M31>*> cp;
    for (u32 c = 0; c < n_cols; c++) {
        cv[c] = to_m31(cols + c * n, n);
        cp.push_back(&cv[c]);
    }
    std::vector<SampleBatch> batches(n_batches);
    for (u32 b = 0; b < n_batches; b++) {
        batches[b].point = qpt(batch_points + 8 * b);
        for (u32 k = batch_start[b]; k < batch_start[b + 1]; k++)
            batches[b].columns_and_values.push_back({col_idx[k], qm(values + 4 * k)});
    }
    sc_to(accumulate_quotients(log_size, cp, qm(random_coeff), batches), out4);
}
u64 orc_grind(const uint8_t* digest, u32 pow_bits) {
    OChannel ch;
    memcpy(ch.digest.b, digest, 32);
    return grind(ch, pow_bits);
}
void orc_prefix_sum(u32* col, u32 log_size) {
    size_t n = (size_t)1 << log_size;
    std::vector<M31> r = inclusive_prefix_sum(to_m31(col, n));
    for (size_t i = 0; i < n; i++) col[i] = r[i].v;
}

// channel: state = 32 digest bytes + u32 n_sent (36 bytes, caller-owned)
static OChannel ch_load(const uint8_t* st) {
    OChannel c;
    memcpy(c.digest.b, st, 32);
    memcpy(&c.n_sent, st + 32, 4);
    return c;
}
static void ch_store(const OChannel& c, uint8_t* st) {
    memcpy(st, c.digest.b, 32);
    memcpy(st + 32, &c.n_sent, 4);
}
void orc_channel_mix_u32s(uint8_t* st, const u32* data, size_t n) {
    OChannel c = ch_load(st);
    c.mix_u32s(data, n);
    ch_store(c, st);
}
void orc_channel_mix_u64(uint8_t* st, u64 v) {
    OChannel c = ch_load(st);
    c.mix_u64(v);
    ch_store(c, st);
}
void orc_channel_draw_secure_felts(uint8_t* st, size_t n, u32* out) {
    OChannel c = ch_load(st);
    auto f = c.draw_secure_felts(n);
    for (size_t i = 0; i < n; i++) f[i].to_u32(out + 4 * i);
    ch_store(c, st);
}
void orc_channel_draw_random_bytes(uint8_t* st, uint8_t* out) {
    OChannel c = ch_load(st);
    Hash h = c.draw_random_bytes();
    memcpy(out, h.b, 32);
    ch_store(c, st);
}

}  // extern "C"

// ------------------------------------------------------------------ whole-proof checkers
#include "host/test_provers.hpp"
#include "oracle_air.hpp"
#include "verifier.hpp"

static int copy_out(const std::vector<uint8_t>& bytes, uint8_t* out, size_t cap, size_t* out_len) {
    if (out_len) *out_len = bytes.size();
    if (out && bytes.size() > cap) return -1;
    if (out) memcpy(out, bytes.data(), bytes.size());
    return 0;
}
static thread_local std::string g_orc_err;

extern "C" {

const char* orc_last_error() { return g_orc_err.c_str(); }

int orc_prove_wide_fibonacci(u32 log_n_rows, u32 n_cols, u32 pow_bits, u32 n_queries, uint8_t* out, size_t cap, size_t* out_len) {
    try {
        cm31::PcsConfig cfg;
        cfg.pow_bits = pow_bits;
        cfg.fri_config.n_queries = n_queries;
        cm31::StarkProof proof = cm31::prove_wide_fibonacci<OracleBackend, OracleComponent<cm31::WideFibonacciEval>>(log_n_rows, n_cols, cfg);
        cm31::ProofWriter w;
        w.proof(proof);
        return copy_out(w.bytes, out, cap, out_len);
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return -2;
    }
}

// verify (wide_fibonacci/mod.rs:214-228): 0 = accepted
int orc_verify_wide_fibonacci(u32 log_n_rows, u32 n_cols, const uint8_t* proof_bytes, size_t len) {
    try {
        cm31::ProofReader r(proof_bytes, len);
        cm31::StarkProof proof = r.proof();
        cm31::RelationSet relations;
        OracleComponent<cm31::WideFibonacciEval> component(cm31::WideFibonacciEval{log_n_rows, n_cols}, &relations);
        cm31::TraceLocationAllocator alloc;
        component.allocate(alloc);
        OChannel channel;
        CommitmentSchemeVerifier cs(proof.config);
        auto sizes = component.trace_log_degree_bounds();
        cs.commit(proof.commitments.at(0), sizes[0], channel);
        cs.commit(proof.commitments.at(1), sizes[1], channel);
        std::vector<const OComponent*> comps = {&component};
        verify(comps, channel, cs, proof);
        return 0;
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return 1;
    }
}

}  // extern "C"

// ------------------------------------------------------------------ cairo-m (fibonacci_loop)
#include "oracle_cairo.hpp"

extern "C" {

// prove_cairo_m on the scalar CPU oracle. timings_ms: preprocessed, trace, interaction, stark, total.
int orc_program_prove(u32 program_id, u32 n, u32 pow_bits, u32 n_queries, uint8_t* out, size_t cap, size_t* out_len, double* timings_ms);
int orc_fib_prove(u32 n, u32 pow_bits, u32 n_queries, uint8_t* out, size_t cap, size_t* out_len, double* timings_ms) {
    return orc_program_prove(cm31::PROGRAM_FIBONACCI_LOOP, n, pow_bits, n_queries, out, cap, out_len, timings_ms);
}
int orc_program_prove(u32 program_id, u32 n, u32 pow_bits, u32 n_queries, uint8_t* out, size_t cap, size_t* out_len, double* timings_ms) {
    try {
        cm31::ProverInput input = cm31::import_from_vm(cm31::run_program(cm31::program_by_id(program_id), n));
        cm31::PcsConfig cfg = cm31::PcsConfig::regular_96_bits();
        cfg.pow_bits = pow_bits;
        cfg.fri_config.n_queries = n_queries;
        cm31::ProveTimings t;
        cm31::CairoProof proof = cm31::prove_cairo_m<OracleAirImpl>(input, cfg, &t);
        if (timings_ms) {
            timings_ms[0] = t.preprocessed_ms;
            timings_ms[1] = t.trace_ms;
            timings_ms[2] = t.interaction_ms;
            timings_ms[3] = t.stark_ms;
            timings_ms[4] = t.total_ms;
        }
        cm31::HostTimer::report();  // CM31_HOST_TIMING=1: host-side sections of the shared protocol driver
        return copy_out(proof.to_bytes(), out, cap, out_len);
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return -2;
    }
}

// One continuation segment (crates/runner/src/vm/mod.rs:158-285) proven on the oracle backend; *n_segments = segments of the run.
int orc_segment_prove(u32 program_id, u32 n, u64 segment_steps, u32 index, u32* n_segments, u32 pow_bits, u32 n_queries, uint8_t* out,
                      size_t cap, size_t* out_len) {
    try {
        std::vector<cm31::VmTrace> segs;
        cm31::VmTrace last = cm31::run_program(cm31::program_by_id(program_id), n, (size_t)1 << 30, &segs, (size_t)segment_steps);
        segs.push_back(std::move(last));
        if (n_segments) *n_segments = (u32)segs.size();
        if (index >= segs.size()) throw std::runtime_error("no such segment");
        cm31::ProverInput input = cm31::import_from_vm(segs[index]);
        cm31::PcsConfig cfg = cm31::PcsConfig::regular_96_bits();
        cfg.pow_bits = pow_bits;
        cfg.fri_config.n_queries = n_queries;
        cm31::CairoProof proof = cm31::prove_cairo_m<OracleAirImpl>(input, cfg, nullptr);
        return copy_out(proof.to_bytes(), out, cap, out_len);
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return -2;
    }
}

int orc_cairo_verify(const uint8_t* proof_bytes, size_t len, u32 pow_bits, u32 n_queries) {
    try {
        cm31::CairoProof proof = cm31::CairoProof::from_bytes(proof_bytes, len, cm31::cairo_component_names());
        cm31::PcsConfig cfg = cm31::PcsConfig::regular_96_bits();
        cfg.pow_bits = pow_bits;
        cfg.fri_config.n_queries = n_queries;
        verify_cairo_m(proof, cfg);
        return 0;
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return 1;
    }
}

// residual of the logup balance (0,0,0,0 expected); also returns fib(n) and the VM step count
int orc_program_logup_residual(u32 program_id, u32 n, const uint8_t* proof_bytes, size_t len, u32* residual_out, u64* info_out);
int orc_fib_logup_residual(u32 n, const uint8_t* proof_bytes, size_t len, u32* residual_out, u64* info_out) {
    return orc_program_logup_residual(cm31::PROGRAM_FIBONACCI_LOOP, n, proof_bytes, len, residual_out, info_out);
}
int orc_program_logup_residual(u32 program_id, u32 n, const uint8_t* proof_bytes, size_t len, u32* residual_out, u64* info_out) {
    try {
        cm31::VmTrace vm = cm31::run_program(cm31::program_by_id(program_id), n);
        cm31::ProverInput input = cm31::import_from_vm(vm);
        cm31::CairoProof proof = cm31::CairoProof::from_bytes(proof_bytes, len, cm31::cairo_component_names());
        logup_residual(proof, input).to_u32(residual_out);
        if (info_out) {
            info_out[0] = vm.return_value;
            info_out[1] = input.n_steps;
            info_out[2] = input.clock_update_data.size();
        }
        return 0;
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return 1;
    }
}

// Poseidon2 permutation of csrc/cairo/poseidon2.hpp; checked against the reference KAT (crates/prover/tests/poseidon2.rs:15-35)
int orc_poseidon2_permutation(const u32* in16, u32* out16) {
    std::array<u32, 16> in;
    for (int i = 0; i < 16; i++) in[i] = in16[i];
    std::array<u32, 16> out = cm31::poseidon2_permutation(in);
    for (int i = 0; i < 16; i++) out16[i] = out[i];
    return 0;
}
// cm31::MemoryModel (csrc/cairo/vm.hpp: the host restatement of adapter::memory::Memory, crates/prover/src/adapter/memory.rs:405-537)
// driven by a script of pushes, so the reference's own Memory::push unit tests (memory.rs:546-858) can be replayed on it.
//   entries  : (address, value[4], clock) x n
//   args_out : MemoryArg per push = (address, prev_clock, clock, prev_val[4], value[4]), 11 words
//   cu_out   : clock_update_data rows (address, prev_clk, value[4]), 6 words each, in push order
//   cells    : present cells of initial_memory then final_memory as (address, value[4], clock, multiplicity), 7 words, ascending address
int orc_memory_push_script(const u32* initial_memory, size_t n_initial, const u32* entries, size_t n, u32* args_out, u32* cu_out,
                           size_t cu_cap, size_t* n_cu, u32* init_cells, u32* final_cells, size_t cells_cap, size_t* n_init_cells,
                           size_t* n_final_cells) {
    using namespace cm31;
    try {
        std::vector<Word4> init(n_initial);
        for (size_t a = 0; a < n_initial; a++)
            for (int k = 0; k < 4; k++) init[a].v[k] = initial_memory[4 * a + k];
        MemoryModel m(init);
        for (size_t i = 0; i < n; i++) {
            const u32* e = entries + 6 * i;
            MemoryModel::Arg a = m.push(e[0], Word4{{e[1], e[2], e[3], e[4]}}, e[5]);
            u32* o = args_out + 11 * i;
            o[0] = a.address, o[1] = a.prev_clock, o[2] = a.clock;
            for (int k = 0; k < 4; k++) o[3 + k] = a.prev_val.v[k], o[7 + k] = a.value.v[k];
        }
        *n_cu = m.clock_update_data.size();
        if (*n_cu > cu_cap) throw std::runtime_error("clock-update capacity");
        for (size_t i = 0; i < *n_cu; i++) {
            const ClockUpdateRow& r = m.clock_update_data[i];
            cu_out[6 * i] = r.address, cu_out[6 * i + 1] = r.prev_clk;
            for (int k = 0; k < 4; k++) cu_out[6 * i + 2 + k] = r.value[k];
        }
        auto dump = [&](const std::vector<MemoryModel::Cell>& cells, u32* out, size_t* n_out) {
            size_t c = 0;
            for (size_t a = 0; a < cells.size(); a++) {
                if (!cells[a].present) continue;
                if (c >= cells_cap) throw std::runtime_error("cell capacity");
                out[7 * c] = (u32)a;
                for (int k = 0; k < 4; k++) out[7 * c + 1 + k] = cells[a].value.v[k];
                out[7 * c + 5] = cells[a].clock, out[7 * c + 6] = cells[a].multiplicity;
                c++;
            }
            *n_out = c;
        };
        dump(m.initial, init_cells, n_init_cells);
        dump(m.final_, final_cells, n_final_cells);
        return 0;
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return 1;
    }
}

// The scenarios of the reference's adapter/merkle.rs tests (:262-424) on build_partial_merkle_tree; 0 = all hold.
int orc_merkle_selftest(void) {
    using namespace cm31;
    try {
        PublicRanges none;
        auto find = [](const std::vector<MerkleNode>& t, u32 index, u32 depth) -> const MerkleNode* {
            for (auto& n : t)
                if (n.index == index && n.depth == depth) return &n;
            return nullptr;
        };
        {  // empty memory: no tree (here: refused)
            std::vector<MerkleNode> t;
            bool threw = false;
            try {
                build_partial_merkle_tree({}, true, none, t);
            } catch (const std::exception&) {
                threw = true;
            }
            if (!threw || !t.empty()) return 1;
        }
        {  // single element: one node per depth 30..1, every parent = hash(left, right), the last parent is the root
            std::vector<MerkleNode> t;
            u32 root = build_partial_merkle_tree({MerkleLeafCell{5, {42, 0, 0, 0}}}, true, none, t);
            if (t.size() != 2 + 29) return 2;  // two leaf pairs at depth 30, then one node per depth 29..1
            for (auto& n : t)
                if (n.parent_value != poseidon2_hash(n.left_value, n.right_value) || n.root != root) return 3;
            if (t.back().depth != 1 || t.back().parent_value != root) return 4;
        }
        {  // two cells: leaves 0..3 and 4..7
            std::vector<MerkleNode> t;
            build_partial_merkle_tree({MerkleLeafCell{0, {10, 11, 12, 13}}, MerkleLeafCell{1, {20, 21, 22, 23}}}, true, none, t);
            const MerkleNode* n = find(t, 0, 30);
            if (!n || n->left_value != 10 || n->right_value != 11) return 5;
            n = find(t, 2, 30);
            if (!n || n->left_value != 12 || n->right_value != 13) return 6;
            n = find(t, 4, 30);
            if (!n || n->left_value != 20 || n->right_value != 21) return 7;
            n = find(t, 0, 29);  // parents of (0,1) and (2,3) meet at depth 29, both real nodes (multiplicity 1)
            if (!n || n->left_multiplicity != 1 || n->right_multiplicity != 1) return 8;
        }
        {  // addresses at both ends: depth range 1..30, missing siblings are default hashes with multiplicity 0
            std::vector<MerkleNode> t;
            build_partial_merkle_tree({MerkleLeafCell{0, {1, 0, 0, 0}}, MerkleLeafCell{(1u << 28) - 1, {2, 0, 0, 0}}}, true, none, t);
            u32 min_depth = 99, max_depth = 0;
            for (auto& n : t) {
                min_depth = std::min(min_depth, n.depth);
                max_depth = std::max(max_depth, n.depth);
            }
            if (min_depth != 1 || max_depth != TREE_HEIGHT) return 9;
            const MerkleNode* n = find(t, 0, 28);
            if (!n || n->right_multiplicity != 0 || n->right_value != poseidon2_default_hashes()[28]) return 10;
        }
        {  // public ranges: multiplicity 2 for program / input cells of the initial tree, output cells of the final tree
            PublicRanges r;
            r.program_start = 0, r.program_end = 1, r.output_start = 7, r.output_end = 8;
            std::vector<MerkleNode> ti, tf;
            std::vector<MerkleLeafCell> cells = {MerkleLeafCell{0, {1, 2, 3, 4}}, MerkleLeafCell{7, {5, 6, 7, 8}}};
            build_partial_merkle_tree(cells, true, r, ti);
            build_partial_merkle_tree(cells, false, r, tf);
            if (find(ti, 0, 30)->left_multiplicity != 2 || find(ti, 28, 30)->left_multiplicity != 1) return 11;
            if (find(tf, 0, 30)->left_multiplicity != 1 || find(tf, 28, 30)->left_multiplicity != 2) return 12;
        }
        return 0;
    } catch (const std::exception& e) {
        g_orc_err = e.what();
        return -1;
    }
}

}  // extern "C"
