// prove_cairo_m: the cairo-m protocol driver, generic over the backend implementation `Impl`
// (CudaAirImpl = product; the test oracle supplies a scalar CPU Impl).
//
// Mirrors crates/prover/src/prover.rs:23-147 (prove_cairo_m), components/mod.rs:94-308
// (Claim::write_trace / InteractionClaim::write_interaction_trace / Relations::draw / Components::new),
// components/opcodes/mod.rs:41-78,223-268 (opcode dispatch + order), public_data.rs:236-412,
// preprocessed/mod.rs:43-82, relations.rs:47.  Transcript order = SURVEY.md Appendix A.
//
// Scope: all 34 components of the reference's statement -- the 26 opcode components,
// memory, merkle, clock_update, poseidon2, range_check_8/16/20 and the bitwise table.
#pragma once
#include <functional>
#include <memory>

#include "../air/cairo_components.hpp"
#include "../host/framework.hpp"
#include "../host/stark.hpp"
#include "vm.hpp"

namespace cm31 {

constexpr u32 PREPROCESSED_TRACE_LOG_SIZE = 20;  // prover.rs:21
constexpr u32 INTERACTION_POW_BITS = 2;          // relations.rs:47
constexpr u32 LOG_N_LANES = 4;                   // minimum component size (store_fp_fp.rs:161)

struct CairoClaim {
    std::vector<std::pair<std::string, u32>> log_sizes;  // component name -> log_size, in mix order
};
struct CairoInteractionClaim {
    std::vector<QM31> claimed_sums;  // same order
};
struct PublicEntry {
    u32 addr;
    u32 value[4];
    u32 clock;
};
struct PublicData {  // public_data.rs:226-266
    Registers initial_registers, final_registers;
    u32 clock;
    u32 initial_root, final_root;
    std::vector<PublicEntry> program, input, output;

    static PublicData from_input(const ProverInput& in) {
        PublicData pd;
        pd.initial_registers = in.initial_registers;
        pd.final_registers = in.final_registers;
        pd.clock = (u32)(in.n_steps % P);
        pd.initial_root = in.initial_root;
        pd.final_root = in.final_root;
        auto extract = [](const std::vector<MemoryRow>& mem, u32 lo, u32 hi, std::vector<PublicEntry>& out) {
            for (const MemoryRow& r : mem)
                if (r.address >= lo && r.address < hi) {
                    PublicEntry e;
                    e.addr = r.address;
                    for (int k = 0; k < 4; k++) e.value[k] = r.value[k];
                    e.clock = r.clock;
                    out.push_back(e);
                }
        };
        extract(in.initial_memory, in.public_ranges.program_start, in.public_ranges.program_end, pd.program);
        extract(in.initial_memory, in.public_ranges.input_start, in.public_ranges.input_end, pd.input);
        extract(in.final_memory, in.public_ranges.output_start, in.public_ranges.output_end, pd.output);
        return pd;
    }
    void mix_into(Blake2sChannel& ch, const PublicRanges& r) const {  // public_data.rs:401-412, :83-139
        u32 head[7] = {initial_registers.pc, initial_registers.fp, final_registers.pc, final_registers.fp, clock, initial_root, final_root};
        ch.mix_u32s(head, 7);
        u32 lens[3] = {r.program_end - r.program_start, r.input_end - r.input_start, r.output_end - r.output_start};
        ch.mix_u32s(lens, 3);
        for (const std::vector<PublicEntry>* v : {&program, &input, &output}) {
            std::vector<u32> w;
            for (const PublicEntry& e : *v) {
                w.push_back(e.addr);
                for (int k = 0; k < 4; k++) w.push_back(e.value[k]);
                w.push_back(e.clock);
            }
            ch.mix_u32s(w.data(), w.size());
        }
    }
};

struct CairoProof {  // crates/prover/src/lib.rs:61-73
    CairoClaim claim;
    CairoInteractionClaim interaction_claim;
    PublicData public_data;
    PublicRanges public_ranges;
    StarkProof stark_proof;
    u64 interaction_pow = 0;

    std::vector<uint8_t> to_bytes() const {
        ProofWriter w;
        write(w);
        return w.bytes;
    }
    void write(ProofWriter& w) const {
        w.u64v(claim.log_sizes.size());
        for (auto& kv : claim.log_sizes) w.u32v(kv.second);
        w.u64v(interaction_claim.claimed_sums.size());
        for (auto& s : interaction_claim.claimed_sums) w.qm(s);
        const PublicData& pd = public_data;
        for (u32 v : {pd.initial_registers.pc, pd.initial_registers.fp, pd.final_registers.pc, pd.final_registers.fp, pd.clock, pd.initial_root, pd.final_root}) w.u32v(v);
        for (u32 v : {public_ranges.program_start, public_ranges.program_end, public_ranges.input_start, public_ranges.input_end,
                      public_ranges.output_start, public_ranges.output_end})
            w.u32v(v);
        for (const std::vector<PublicEntry>* v : {&pd.program, &pd.input, &pd.output}) {
            w.u64v(v->size());
            for (const PublicEntry& e : *v) {
                w.u32v(e.addr);
                for (int k = 0; k < 4; k++) w.u32v(e.value[k]);
                w.u32v(e.clock);
            }
        }
        w.u64v(interaction_pow);
        w.proof(stark_proof);
    }
    static CairoProof from_bytes(const uint8_t* data, size_t len, const std::vector<std::string>& component_names) {
        ProofReader r(data, len);
        CairoProof p;
        u64 n = r.u64v();
        if (n != component_names.size()) throw std::runtime_error("proof: component count mismatch");
        for (u64 i = 0; i < n; i++) p.claim.log_sizes.push_back({component_names[i], r.u32v()});
        u64 m = r.u64v();
        for (u64 i = 0; i < m; i++) p.interaction_claim.claimed_sums.push_back(r.qm());
        PublicData& pd = p.public_data;
        pd.initial_registers.pc = r.u32v();
        pd.initial_registers.fp = r.u32v();
        pd.final_registers.pc = r.u32v();
        pd.final_registers.fp = r.u32v();
        pd.clock = r.u32v();
        pd.initial_root = r.u32v();
        pd.final_root = r.u32v();
        p.public_ranges.program_start = r.u32v();
        p.public_ranges.program_end = r.u32v();
        p.public_ranges.input_start = r.u32v();
        p.public_ranges.input_end = r.u32v();
        p.public_ranges.output_start = r.u32v();
        p.public_ranges.output_end = r.u32v();
        for (std::vector<PublicEntry>* v : {&pd.program, &pd.input, &pd.output}) {
            u64 k = r.u64v();
            for (u64 i = 0; i < k; i++) {
                PublicEntry e;
                e.addr = r.u32v();
                for (int j = 0; j < 4; j++) e.value[j] = r.u32v();
                e.clock = r.u32v();
                v->push_back(e);
            }
        }
        p.interaction_pow = r.u64v();
        ProofReader rest(data + r.pos, len - r.pos);
        p.stark_proof = rest.proof();
        return p;
    }
};

inline RelationSet draw_cairo_relations(Blake2sChannel& ch) {  // components/mod.rs:311-323
    RelationSet rs;
    for (int r = 0; r < N_CAIRO_RELATIONS; r++) rs.relations.push_back(RelationElements::draw(ch, cairo_relation_size(r)));
    return rs;
}

// PreProcessedTraceBuilder::default (preprocessed/mod.rs:75-82): bitwise x4, rc8, rc16, rc20
inline std::vector<std::string> cairo_preprocessed_ids() {
    return {BitwiseEval::column_id(0), BitwiseEval::column_id(1), BitwiseEval::column_id(2), BitwiseEval::column_id(3),
            "range_check_8", "range_check_16", "range_check_20"};
}
inline std::vector<u32> cairo_preprocessed_log_sizes() {
    return {BITWISE_STACKED_LOG_SIZE, BITWISE_STACKED_LOG_SIZE, BITWISE_STACKED_LOG_SIZE, BITWISE_STACKED_LOG_SIZE, 8, 16, 20};
}
inline std::vector<std::string> cairo_component_names() {
    std::vector<std::string> names;
#define CM31_X(E) names.push_back(E::name());
    CM31_OPCODE_EVALS(CM31_X)
#undef CM31_X
    // components/mod.rs:77-104: opcodes, memory, merkle, clock_update, poseidon2, range checks, bitwise
    for (const char* n : {"memory", "merkle", "clock_update", "poseidon2", "range_check_8", "range_check_16", "range_check_20", "bitwise"}) names.push_back(n);
    return names;
}
inline size_t n_opcode_components() {
    size_t n = 0;
#define CM31_X(E) n++;
    CM31_OPCODE_EVALS(CM31_X)
#undef CM31_X
    return n;
}

inline u32 padded_log_size(size_t n_real) {
    u32 l = LOG_N_LANES;
    while (((size_t)1 << l) < n_real) l++;
    return l;
}

// The full component set; shared by prover and verifier.
template <class Impl>
struct CairoComponents {
    typedef typename Impl::B B;
    template <class Eval>
    using Comp = typename Impl::template Component<Eval>;
#define CM31_X(E) std::unique_ptr<Comp<E>> c_##E;
    CM31_OPCODE_EVALS(CM31_X)
#undef CM31_X
    std::unique_ptr<Comp<MemoryEval>> memory;
    std::unique_ptr<Comp<MerkleEval>> merkle;
    std::unique_ptr<Comp<ClockUpdateEval>> clock_update;
    std::unique_ptr<Comp<Poseidon2Eval>> poseidon2;
    std::unique_ptr<Comp<RangeCheckEval>> rc8, rc16, rc20;
    std::unique_ptr<Comp<BitwiseEval>> bitwise;

    // `log_sizes` in cairo_component_names() order
    CairoComponents(const std::vector<u32>& ls, const RelationSet* rel) {
        auto base = [](u32 l) {
            OpcodeEvalBase b;
            b.log_size_ = l;
            return b;
        };
        size_t i = 0;
#define CM31_X(E) c_##E.reset(new Comp<E>(E{base(ls.at(i++))}, rel));
        CM31_OPCODE_EVALS(CM31_X)
#undef CM31_X
        memory.reset(new Comp<MemoryEval>(MemoryEval{base(ls.at(i++))}, rel));
        merkle.reset(new Comp<MerkleEval>(MerkleEval{base(ls.at(i++))}, rel));
        clock_update.reset(new Comp<ClockUpdateEval>(ClockUpdateEval{base(ls.at(i++))}, rel));
        poseidon2.reset(new Comp<Poseidon2Eval>(Poseidon2Eval{base(ls.at(i++))}, rel));
        rc8.reset(new Comp<RangeCheckEval>(RangeCheckEval{base(ls.at(i)), REL_RC8}, rel));
        rc16.reset(new Comp<RangeCheckEval>(RangeCheckEval{base(ls.at(i + 1)), REL_RC16}, rel));
        rc20.reset(new Comp<RangeCheckEval>(RangeCheckEval{base(ls.at(i + 2)), REL_RC20}, rel));
        bitwise.reset(new Comp<BitwiseEval>(BitwiseEval{base(ls.at(i + 3))}, rel));
    }
    template <class Fn>
    void for_each(Fn fn) {
#define CM31_X(E) fn(*c_##E);
        CM31_OPCODE_EVALS(CM31_X)
#undef CM31_X
        fn(*memory);
        fn(*merkle);
        fn(*clock_update);
        fn(*poseidon2);
        fn(*rc8);
        fn(*rc16);
        fn(*rc20);
        fn(*bitwise);
    }
    void allocate(TraceLocationAllocator& alloc) {
        for_each([&](auto& c) { c.allocate(alloc); });
    }
    std::vector<const ComponentProver<B>*> provers() {
        std::vector<const ComponentProver<B>*> v;
        for_each([&](auto& c) { v.push_back(&c); });
        return v;
    }
    void set_claimed_sums(const std::vector<QM31>& sums) {
        size_t i = 0;
        for_each([&](auto& c) { c.claimed_sum = sums.at(i++); });
    }
};

// Per-phase wall-clock (host) milliseconds, filled when non-null (bench.py reports them).
struct ProveTimings {
    double preprocessed_ms = 0, trace_ms = 0, interaction_ms = 0, stark_ms = 0, total_ms = 0;
};

// Prover input resident in backend memory (device buffers for the CUDA path): the AoS records of
// ProverInput flattened to u32 words, per component, in component order.  stage_input is the
// host->device copy of a proof's inputs; prove_cairo_m itself never touches the host vectors.
template <class Impl>
struct StagedInput {
    struct Rows {
        typename Impl::Words words;
        size_t n_real = 0;
        u32 mark = 0;  // staging mark: the rows have landed once it is reached
    };
    typename Impl::Words accesses;
    size_t n_accesses = 0;
    u32 accesses_mark = 0;
    std::vector<Rows> opcode;  // one per opcode component, CM31_OPCODE_EVALS order
    Rows memory, merkle, clock_update, poseidon2;
    size_t bytes = 0;  // total bytes staged
    uint64_t bg_ticket = 0;  // deferred staging (cm31_bg_defer): the copies are issued by cm31_bg_release(bg_ticket) at the latest
};

// The O(memory footprint) tables of a prover input: boundary-memory rows, Merkle nodes and the Poseidon2 states derived from
// them.  Shared by stage_input (host adapter output) and the device adapter (cm31_adapter_import), whose per-step tables are
// produced in HBM directly.
template <class Impl>
void stage_boundary_rows(const ProverInput& input, StagedInput<Impl>& st) {
    {  // memory (components/memory.rs:93-195): initial rows then final rows
        std::vector<u32> rows;
        for (const std::vector<MemoryRow>* v : {&input.initial_memory, &input.final_memory})
            for (const MemoryRow& r : *v)
                for (u32 w : {r.address, r.clock, r.value[0], r.value[1], r.value[2], r.value[3], r.multiplicity, r.root}) rows.push_back(w);
        st.memory.n_real = input.initial_memory.size() + input.final_memory.size();
        st.memory.words = Impl::upload_words(rows.data(), rows.size());
        st.bytes += rows.size() * 4;
    }
    {  // merkle (components/merkle.rs:82-104: initial tree then final tree, each node with its tree's root) and the
       // poseidon2 inputs derived from the same nodes (adapter/mod.rs:165-174: state = [left, right, 0, ..])
        std::vector<u32> rows, states;
        for (const MerkleNode& nd : input.merkle_nodes) {
            for (u32 w : {nd.index, nd.depth, nd.left_value, nd.right_value, nd.parent_value, nd.left_multiplicity, nd.right_multiplicity,
                          nd.parent_multiplicity, nd.root})
                rows.push_back(w);
            states.push_back(nd.left_value);
            states.push_back(nd.right_value);
            for (int k = 2; k < POSEIDON2_T; k++) states.push_back(0);
        }
        st.merkle.n_real = st.poseidon2.n_real = input.merkle_nodes.size();
        st.merkle.words = Impl::upload_words(rows.data(), rows.size());
        st.poseidon2.words = Impl::upload_words(states.data(), states.size());
        st.bytes += (rows.size() + states.size()) * 4;
    }
}

template <class Impl>
StagedInput<Impl> stage_input(const ProverInput& input, StagedInput<Impl>* recycle = nullptr) {
    StagedInput<Impl> st;
    // pass 1: every destination buffer (stream-ordered allocations on the proof's stream); pass 2, after ONE ordering
    // point (Impl::staging_begin), the copies on the background stream, each followed by its mark.
    // `recycle`: a consumed StagedInput of the same shape (the previous segment's slot) whose big buffers are taken over
    // instead of allocated — a prefetch issued while a proof is in flight must not compete with that proof for the
    // allocator's cached blocks (measured: the pool kept growing and a pipelined step took up to 2x a serial one).
    st.n_accesses = input.data_accesses.size();
    if (recycle && recycle->n_accesses == st.n_accesses && recycle->accesses.size() >= std::max<size_t>(4, st.n_accesses * 4))
        st.accesses = std::move(recycle->accesses);
    else
        st.accesses = Impl::alloc_words(st.n_accesses * 4);
    st.bytes += st.n_accesses * 16;
    std::vector<std::vector<const std::vector<Bundle>*>> parts_of;
    auto opcode_rows = [&](const std::vector<u32>& opcodes) {
        typename StagedInput<Impl>::Rows r;
        std::vector<const std::vector<Bundle>*> parts;
        for (u32 op : opcodes) {
            auto it = input.states_by_opcodes.find(op);
            if (it != input.states_by_opcodes.end() && !it->second.empty()) parts.push_back(&it->second);
        }
        for (auto* v : parts) r.n_real += v->size();
        size_t c = st.opcode.size();
        if (recycle && c < recycle->opcode.size() && recycle->opcode[c].n_real == r.n_real &&
            recycle->opcode[c].words.size() >= std::max<size_t>(4, r.n_real * 12))
            r.words = std::move(recycle->opcode[c].words);
        else
            r.words = Impl::alloc_words(r.n_real * 12);
        st.bytes += r.n_real * sizeof(Bundle);
        st.opcode.push_back(std::move(r));
        parts_of.push_back(std::move(parts));
    };
#define CM31_X(E) opcode_rows(E::opcodes());
    CM31_OPCODE_EVALS(CM31_X)
#undef CM31_X
    Impl::staging_begin();
    Impl::copy_words(st.accesses, 0, (const u32*)input.data_accesses.data(), st.n_accesses * 4);
    st.accesses_mark = Impl::staging_mark();
    for (size_t c = 0; c < st.opcode.size(); c++) {
        size_t at = 0;
        for (auto* v : parts_of[c]) {  // straight from the adapter's (page-locked) vectors, no host-side concatenation
            Impl::copy_words(st.opcode[c].words, at * 12, (const u32*)v->data(), v->size() * 12);
            at += v->size();
        }
        st.opcode[c].mark = Impl::staging_mark();
    }
    stage_boundary_rows<Impl>(input, st);
    {  // clock_update (components/clock_update.rs:70-160)
        std::vector<u32> rows;
        for (const ClockUpdateRow& r : input.clock_update_data)
            for (u32 w : {r.address, r.prev_clk, r.value[0], r.value[1], r.value[2], r.value[3]}) rows.push_back(w);
        st.clock_update.n_real = input.clock_update_data.size();
        st.clock_update.words = Impl::upload_words(rows.data(), rows.size());
        st.bytes += rows.size() * 4;
    }
    return st;
}

// Longest-processing-time-first assignment of components to `world` ranks (deterministic: ties go to the lower index / rank).
inline std::vector<int> assign_component_owners(const std::vector<double>& cost, int world) {
    std::vector<int> owner(cost.size(), 0);
    if (world <= 1) return owner;
    std::vector<size_t> order(cost.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cost[a] > cost[b]; });
    std::vector<double> load((size_t)world, 0.0);
    for (size_t i : order) {
        int best = 0;
        for (int r = 1; r < world; r++)
            if (load[(size_t)r] < load[(size_t)best]) best = r;
        owner[i] = best;
        load[(size_t)best] += cost[i];
    }
    return owner;
}

template <class Impl>
CairoProof prove_cairo_m(const ProverInput& input, const StagedInput<Impl>& staged, PcsConfig pcs_config, ProveTimings* timings = nullptr) {
    typedef typename Impl::B B;
    typedef typename B::Col Col;
    auto t0 = Impl::now_ms();
    Span<B> whole_span("prove_cairo_m");  // P/src/prover.rs:30
    B::shard_begin_proof();  // single-proof sharding: arena reset + barrier; a no-op otherwise
    struct ShardEnd {
        ~ShardEnd() { B::shard_end_proof(); }
    } shard_end;
    Impl::staging_release_point(0);  // a prefetch recorded for the NEXT segment starts to travel (throttled) under this proof
    Blake2sChannel channel;
    pcs_config.mix_into(channel);

    // Component ownership when the proof is sharded over the GPUs of a node: all columns of a component stay on one rank
    // (its constraint evaluation and logup generation are local); components are dealt out longest-processing-time first
    // by padded rows x trace columns.  Every rank computes the same assignment.
    {
        std::vector<double> cost;
        size_t oi = 0;
#define CM31_X(E) cost.push_back((double)((size_t)1 << padded_log_size(staged.opcode.at(oi++).n_real)) * (double)E::N_TRACE_COLUMNS);
        CM31_OPCODE_EVALS(CM31_X)
#undef CM31_X
        cost.push_back((double)((size_t)1 << padded_log_size(staged.memory.n_real)) * MemoryEval::N_TRACE_COLUMNS);
        cost.push_back((double)((size_t)1 << padded_log_size(staged.merkle.n_real)) * MerkleEval::N_TRACE_COLUMNS);
        cost.push_back((double)((size_t)1 << padded_log_size(staged.clock_update.n_real)) * ClockUpdateEval::N_TRACE_COLUMNS);
        cost.push_back((double)((size_t)1 << padded_log_size(staged.poseidon2.n_real)) * Poseidon2Eval::N_TRACE_COLUMNS);
        for (u32 bits : {8u, 16u, 20u, (u32)BITWISE_STACKED_LOG_SIZE}) cost.push_back((double)((size_t)1 << bits) * 2.0);
        B::set_component_owners(assign_component_owners(cost, B::shard_world()));
    }
    size_t scope_index = 0;  // claim-order index of the component whose trace is being written

    // trace_log_size (prover.rs:38-53)
    size_t max_rows = 1;
    for (auto& r : staged.opcode) max_rows = std::max(max_rows, r.n_real);
    u32 trace_log_size = std::max(PREPROCESSED_TRACE_LOG_SIZE, padded_log_size(max_rows));
    trace_log_size = std::max(trace_log_size, padded_log_size(staged.memory.n_real));
    trace_log_size = std::max(trace_log_size, padded_log_size(staged.merkle.n_real));
    typename B::Twiddles twiddles;
    B::precompute_twiddles(trace_log_size + pcs_config.fri_config.log_blowup_factor + 2, twiddles);
    CommitmentSchemeProver<B> commitment_scheme(pcs_config, &twiddles);

    CairoProof proof;
    proof.public_data = PublicData::from_input(input);
    proof.public_ranges = input.public_ranges;
    proof.public_data.mix_into(channel, input.public_ranges);

    // ---- tree 0: preprocessed (stacked bitwise table x4, range_check_8/16/20 value columns)
    std::vector<u32> rc_bits = {8, 16, 20};
    // lookup tables with a multiplicity component, in component order: (relation, log2 rows)
    const std::vector<std::pair<int, u32>> tables = {{REL_RC8, 8}, {REL_RC16, 16}, {REL_RC20, 20}, {REL_BITWISE, BITWISE_STACKED_LOG_SIZE}};
    auto preprocessed_columns = [&]() {
        std::vector<Col> cols;
        for (int k = 0; k < 4; k++) cols.push_back(Impl::bitwise_table_col(k));
        for (u32 bits : rc_bits) cols.push_back(Impl::iota((size_t)1 << bits));
        return cols;
    };
    {
        std::vector<CircleEvaluation<B>> pre;
        std::vector<u32> sizes = cairo_preprocessed_log_sizes();
        std::vector<Col> cols = preprocessed_columns();
        for (size_t i = 0; i < cols.size(); i++) pre.push_back(CircleEvaluation<B>{std::move(cols[i]), sizes[i]});
        commitment_scheme.commit_evals(std::move(pre), channel);
    }
    auto t1 = Impl::now_ms();

    // ---- tree 1: execution traces
    // the access log (copied first) must be complete before any trace fill; each component then waits for its own rows only
    Impl::staging_wait(staged.accesses_mark);
    std::vector<u32> log_sizes;
    std::vector<std::vector<CircleEvaluation<B>>> traces;  // per component (kept until tree 2 is built)
    size_t opcode_index = 0;
    std::vector<Col> padding_inputs;  // shared by the components without real rows; alive until their batch has been issued
    auto opcode_trace = [&](auto eval_tag) {
        typedef decltype(eval_tag) Eval;
        const auto& rows = staged.opcode.at(opcode_index++);
        u32 ls = padded_log_size(rows.n_real);
        B::component_scope_index(scope_index++);  // sharded proof: only the owner fills this component's trace
        Eval eval;
        eval.log_size_ = ls;
        log_sizes.push_back(ls);
        if (rows.n_real == 0 && Impl::batch_small_components()) {
            // an unused opcode component: 16 rows of ExecutionBundle::default().  Its inputs are the same columns for every
            // such component and its trace program is recorded, not launched (one batched launch for all of them below)
            B::lane(ls);  // the side lane: recorded programs are issued there too, so their columns are allocated, written and
                          // (for the shared inputs) read in one stream order
            if (padding_inputs.empty()) padding_inputs = Impl::unpack_padding(staged.accesses, staged.n_accesses, ls);
            typename Impl::BatchScope batch(true);
            traces.push_back(Impl::template write_trace<Eval>(eval, padding_inputs, 0));
            return;
        }
        B::lane(ls);  // small components go to the side lane; every temporary below dies on the lane that used it
        Impl::staging_wait(rows.mark);
        std::vector<Col> inputs = Impl::unpack_bundles(rows.words, rows.n_real, staged.accesses, staged.n_accesses, ls);
        traces.push_back(Impl::template write_trace<Eval>(eval, inputs, (u32)rows.n_real));
    };
    B::prepare();
    static const bool phase_debug = getenv("CM31_PHASE_DEBUG") != nullptr;
    double dbg_t = Impl::now_ms();
    auto dbg = [&](const char* what) {
        if (!phase_debug) return;
        double now = Impl::now_ms();
        fprintf(stderr, "[phase] %-28s %8.3f ms\n", what, now - dbg_t);
        dbg_t = now;
    };
#define CM31_X(E) opcode_trace(E{}); dbg(#E);
    CM31_OPCODE_EVALS(CM31_X)
#undef CM31_X
    {
        u32 ls = padded_log_size(staged.memory.n_real);
        B::component_scope_index(scope_index++);
        B::lane(ls);
        std::vector<Col> inputs = Impl::unpack_rows(staged.memory.words, staged.memory.n_real, 8, ls);
        MemoryEval eval;
        eval.log_size_ = ls;
        log_sizes.push_back(ls);
        traces.push_back(Impl::template write_trace<MemoryEval>(eval, inputs, (u32)staged.memory.n_real));
    }
    {
        u32 ls = padded_log_size(staged.merkle.n_real);
        B::component_scope_index(scope_index++);
        B::lane(ls);
        std::vector<Col> inputs = Impl::unpack_rows(staged.merkle.words, staged.merkle.n_real, 9, ls);
        MerkleEval eval;
        eval.log_size_ = ls;
        log_sizes.push_back(ls);
        traces.push_back(Impl::template write_trace<MerkleEval>(eval, inputs, (u32)staged.merkle.n_real));
    }
    {
        u32 ls = padded_log_size(staged.clock_update.n_real);
        B::component_scope_index(scope_index++);
        B::lane(ls);
        std::vector<Col> inputs = Impl::unpack_rows(staged.clock_update.words, staged.clock_update.n_real, 6, ls);
        ClockUpdateEval eval;
        eval.log_size_ = ls;
        log_sizes.push_back(ls);
        traces.push_back(Impl::template write_trace<ClockUpdateEval>(eval, inputs, (u32)staged.clock_update.n_real));
    }
    {
        u32 ls = padded_log_size(staged.poseidon2.n_real);
        B::component_scope_index(scope_index++);
        B::lane(ls);
        std::vector<Col> inputs = Impl::unpack_rows(staged.poseidon2.words, staged.poseidon2.n_real, POSEIDON2_T, ls);
        Poseidon2Eval eval;
        eval.log_size_ = ls;
        log_sizes.push_back(ls);
        traces.push_back(Impl::template write_trace<Poseidon2Eval>(eval, inputs, (u32)staged.poseidon2.n_real));
    }
    B::component_scope(-1);
    dbg("memory..poseidon2 traces");
    Impl::air_batch_flush();  // the trace programs of the unused components: one launch
    dbg("trace batch flush");
    B::lanes_join();
    padding_inputs.clear();
    dbg("join");
    // range-check multiplicities: histogram of every value the opcode components look up
    // (opcodes/mod.rs:83-105 providers; range_check_macro.rs:72-84).  The AIR graphs drive it.
    RelationSet dummy_relations;
    for (int r = 0; r < N_CAIRO_RELATIONS; r++) dummy_relations.relations.push_back(RelationElements::dummy(cairo_relation_size(r)));
    {
        std::vector<u32> ls_all = log_sizes;
        for (auto& tb : tables) ls_all.push_back(tb.second);
        CairoComponents<Impl> shape(ls_all, &dummy_relations);
        // all bins are zeroed on lane 0 before the first histogram kernel forks to the side lane; the histograms of
        // different components only meet in atomicAdds
        std::vector<Col> all_bins;
        for (auto& tb : tables) all_bins.push_back(B::zeros((size_t)1 << tb.second));
        typename Impl::BatchScope lookup_batch(true);  // lookups of the 16-row components: recorded, one launch below
        for (size_t ti = 0; ti < tables.size(); ti++) {
            auto& tb = tables[ti];
            Col& bins = all_bins[ti];
            size_t ci = 0;
            auto emit = [&](auto& comp) {
                if (ci < n_opcode_components()) {  // opcode components only
                    std::vector<const Col*> tc;
                    for (auto& e : traces[ci]) tc.push_back(&e.values);
                    B::component_scope_index(ci);  // sharded proof: the owner counts its component's lookups
                    B::lane(log_sizes[ci]);
                    Impl::emit_lookups(comp, tb.first, tc, bins);
                }
                ci++;
            };
            shape.for_each(emit);
        }
        B::component_scope(-1);
        dbg("lookups issued");
        Impl::air_batch_flush();
        B::lanes_join();
        dbg("lookups flush+join");
        for (auto& bins : all_bins) B::allreduce_bins(bins);  // sharded proof: multiplicities are sums over the ranks' components
        for (size_t ti = 0; ti < tables.size(); ti++) {
            auto& tb = tables[ti];
            log_sizes.push_back(tb.second);
            std::vector<CircleEvaluation<B>> t;
            t.push_back(CircleEvaluation<B>{std::move(all_bins[ti]), tb.second});
            traces.push_back(std::move(t));
        }
    }
    std::vector<std::string> names = cairo_component_names();
    for (size_t i = 0; i < names.size(); i++) {
        proof.claim.log_sizes.push_back({names[i], log_sizes[i]});
        channel.mix_u64(log_sizes[i]);  // Claim::mix_into
    }
    // Tree 1 borrows the trace columns (out-of-place interpolation): the interaction trace needs the
    // trace-domain values again for the logup programs.
    Impl::idle_gate_open();  // a previous proof's deferred tail (host work) runs while this commitment's kernels execute
    Impl::staging_release_point(1);  // the next segment's input may start to travel now: long FFT / Merkle kernels follow
    {
        std::vector<const CircleEvaluation<B>*> all;
        for (auto& comp : traces)
            for (auto& e : comp) all.push_back(&e);
        commitment_scheme.commit_evals_keep(all, channel);
    }
    // "lookup outside its table": a witness value that no range-check / bitwise table holds raised the device error word
    // in the histogram kernels above; it is read here, where the root of tree 1 has just synchronised the host anyway
    dbg("tree 1 commit");
    Impl::check_lookups();
    auto t2 = Impl::now_ms();

    // ---- interaction: PoW, relations, logup columns (prover.rs:84-102)
    proof.interaction_pow = B::grind(channel.digest(), INTERACTION_POW_BITS);
    channel.mix_u64(proof.interaction_pow);
    RelationSet relations = draw_cairo_relations(channel);
    CairoComponents<Impl> components(log_sizes, &relations);
    {
        std::vector<CircleEvaluation<B>> interaction;
        std::vector<Col> pre_cols = preprocessed_columns();
        auto pre_lookup = [&](const std::string& id) -> const Col* {
            std::vector<std::string> ids = cairo_preprocessed_ids();
            for (size_t i = 0; i < ids.size(); i++)
                if (ids[i] == id) return &pre_cols[i];
            throw std::logic_error("unknown preprocessed column " + id);
        };
        size_t ci = 0;
        {
            typename Impl::BatchScope logup_batch(true);  // logup programs + finalisations of the 16-row components: two launches
            components.for_each([&](auto& comp) {
                std::vector<const Col*> tc;
                for (auto& e : traces[ci]) tc.push_back(&e.values);
                B::component_scope_index(ci);  // sharded proof: the owner generates this component's logup columns
                B::lane(comp.log_size());
                auto cols = comp.gen_interaction_trace(tc, pre_lookup);
                for (auto& e : cols) interaction.push_back(std::move(e));
                ci++;
            });
        }
        B::component_scope(-1);
        Impl::air_batch_flush();
        B::lanes_join();
        Impl::collect_claimed_sums(components);
        components.for_each([&](auto& comp) { proof.interaction_claim.claimed_sums.push_back(comp.claimed_sum); });
        for (auto& s : proof.interaction_claim.claimed_sums) channel.mix_felts({s});  // InteractionClaim::mix_into
        traces.clear();
        Impl::idle_gate_open();  // second stage of a previous proof's deferred tail (assembly + serialisation)
        Impl::staging_release_point(2);
        commitment_scheme.commit_evals(std::move(interaction), channel);
    }
    auto t3 = Impl::now_ms();

    // ---- STARK (prover.rs:104-131)
    TraceLocationAllocator alloc(cairo_preprocessed_ids());
    components.allocate(alloc);
    Impl::staging_release_point(3);
    proof.stark_proof = prove<B>(components.provers(), channel, commitment_scheme);
    auto t4 = Impl::now_ms();
    if (timings) {
        timings->preprocessed_ms = t1 - t0;
        timings->trace_ms = t2 - t1;
        timings->interaction_ms = t3 - t2;
        timings->stark_ms = t4 - t3;
        timings->total_ms = t4 - t0;
    }
    return proof;
}

// Host-input form (the reference's `prove_cairo_m(&mut ProverInput, ..)`): stages, then proves.
template <class Impl>
CairoProof prove_cairo_m(const ProverInput& input, PcsConfig pcs_config, ProveTimings* timings = nullptr) {
    StagedInput<Impl> staged = stage_input<Impl>(input);
    return prove_cairo_m<Impl>(input, staged, pcs_config, timings);
}

}  // namespace cm31
