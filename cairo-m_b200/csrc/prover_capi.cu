// Proof-level entry points of libcm31 (host protocol driver over the CUDA backend ops).
#include <cstring>

#include "common.cuh"
#include "host/cuda_backend.hpp"
#include "host/framework.hpp"
#include "host/test_provers.hpp"

using namespace cm31;

static int write_out(const std::vector<uint8_t>& bytes, uint8_t* out, size_t cap, size_t* out_len) {
    if (out_len) *out_len = bytes.size();
    CM_REQUIRE(out == nullptr || bytes.size() <= cap, "proof buffer too small");
    if (out) memcpy(out, bytes.data(), bytes.size());
    return 0;
}

extern "C" {

int cm31_test_prove_wide_fibonacci(uint32_t log_n_rows, uint32_t n_cols, uint32_t pow_bits, uint32_t n_queries,
                              uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    try {
        PcsConfig cfg;
        cfg.pow_bits = pow_bits;
        cfg.fri_config.n_queries = n_queries;
        StarkProof proof = prove_wide_fibonacci<CudaBackend, FrameworkComponent<CudaBackend, WideFibonacciEval>>(log_n_rows, n_cols, cfg);
        ProofWriter w;
        w.proof(proof);
        return write_out(w.bytes, proof_out, proof_cap, proof_len);
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}

}  // extern "C"

// ------------------------------------------------------------------ cairo-m
#include <deque>

#include "cairo/prover.hpp"
#include "cairo/proof_json.hpp"
#include "host/cuda_air_impl.hpp"

struct cm31_prover_input {
    cm31::ProverInput input;
    uint32_t return_value = 0;
    std::unique_ptr<cm31::StagedInput<cm31::CudaAirImpl>> staged;  // set by cm31_input_upload
    std::deque<std::unique_ptr<cm31::StagedInput<cm31::CudaAirImpl>>> prefetched;  // cm31_input_prefetch, consumed oldest first
    std::unique_ptr<cm31::StagedInput<cm31::CudaAirImpl>> spare;   // the last consumed prefetch slot: its buffers take the next upload
    std::vector<void*> pinned;                                     // host ranges registered with CUDA
    std::vector<uint32_t> desc_ids, desc_bundles;  // cm31_input_describe: states_by_opcodes flattened
    std::vector<uint64_t> desc_start;
    bool device_adapted = false;  // cm31_adapter_import: the per-step tables exist in HBM only (`staged`), not in `input`
    uint64_t adapted_info[2] = {0, 0};  // data accesses, bytes of runner logs uploaded
    // prefetches that were recorded but never released (deferred staging) must not be issued into freed buffers
    void drop_prefetched() {
        for (auto& st : prefetched)
            if (st && st->bg_ticket) cm31_bg_cancel(st->bg_ticket);
        prefetched.clear();
    }
    ~cm31_prover_input() {
        drop_prefetched();
        for (void* p : pinned) cudaHostUnregister(p);
    }
};

static void pin_range(cm31_prover_input* h, const void* p, size_t bytes) {
    // page-lock the adapter's big vectors so the per-proof H2D copies run at PCIe/C2C speed;
    // best effort (no device here, or the range is tiny): a failure only means a slower copy.
    if (bytes < (1u << 16)) return;
    if (cudaHostRegister((void*)p, bytes, cudaHostRegisterDefault) == cudaSuccess) h->pinned.push_back((void*)p);
    else cudaGetLastError();
}

extern "C" {

// Runs the (host) VM + adapter for fibonacci_loop(n): the prover input of crates/prover/src/adapter.
int cm31_test_program_input_create(uint32_t program_id, uint32_t n, cm31_prover_input** out);
int cm31_test_fib_input_create(uint32_t n, cm31_prover_input** out) { return cm31_test_program_input_create(PROGRAM_FIBONACCI_LOOP, n, out); }
// program_id: 0 = fibonacci_loop(n), 1 = array_sum(n) (call / frame-pointer / double-deref / assert / le opcodes), 2 = u32_counter(n), 3 = u32_mix(n) (u32 mul / divrem / eq / lt and the two-word *_fp_imm u32 instructions)
int cm31_test_program_input_create(uint32_t program_id, uint32_t n, cm31_prover_input** out) {
    try {
        CM_REQUIRE(out != nullptr, "program_input_create: null out");
        VmTrace vm = run_program(program_by_id(program_id), n);
        cm31_prover_input* h = new cm31_prover_input();
        h->input = import_from_vm(vm);
        h->return_value = vm.return_value;
        pin_range(h, h->input.data_accesses.data(), h->input.data_accesses.size() * sizeof(DataAccess));
        for (auto& kv : h->input.states_by_opcodes) pin_range(h, kv.second.data(), kv.second.size() * sizeof(Bundle));
        *out = h;
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
// Test hook: corrupt one value of the adapter output so that a POLYNOMIAL constraint of the store_fp_fp
// AIR (dst_val - res, store_fp_fp.rs evaluate) no longer holds; the prover must then refuse
// (ConstraintsNotSatisfied, stwo prover/mod.rs:76-82).  A corruption that only unbalances a lookup would
// still prove — as in the reference — and be caught by the verifier's logup-sum check instead.
//   kind 0: the value written by the middle StoreAddFpFp step;  kind 1: the second operand read by the middle StoreSubFpFp step
int cm31_test_input_tamper(cm31_prover_input* h, uint32_t kind) {
    CM_REQUIRE(h != nullptr, "input_tamper: null handle");
    CM_REQUIRE(!h->device_adapted, "input_tamper: not available on a device-adapted input");
    h->staged.reset();
    h->drop_prefetched();
    auto it = h->input.states_by_opcodes.find(kind == 0 ? OP_STORE_ADD_FP_FP : OP_STORE_SUB_FP_FP);
    CM_REQUIRE(it != h->input.states_by_opcodes.end() && !it->second.empty(), "input_tamper: the program has no such step");
    const Bundle& b = it->second[it->second.size() / 2];
    CM_REQUIRE(b.span_len == 3, "input_tamper: unexpected access span");
    DataAccess& a = h->input.data_accesses[b.span_start + (kind == 0 ? 2 : 1)];
    a.value = m31_add(a.value, 1);
    return 0;
}
int cm31_input_destroy(cm31_prover_input* h) {
    delete h;
    return 0;
}
// Copies the prover input to HBM once; later cm31_prove_cairo_m calls on this handle skip the
// host->device copy (bench.py's device-resident `value`; without it every proof stages its input).
int cm31_input_upload(cm31_prover_input* h) {
    try {
        CM_REQUIRE(h != nullptr, "input_upload: null handle");
        if (h->device_adapted) return 0;  // already resident
        h->staged.reset(new StagedInput<CudaAirImpl>(stage_input<CudaAirImpl>(h->input)));
        return cm31_sync();
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
int cm31_input_release_device(cm31_prover_input* h) {
    CM_REQUIRE(h != nullptr, "input_release_device: null handle");
    CM_REQUIRE(!h->device_adapted, "input_release_device: a device-adapted input has no host copy to fall back to");
    h->staged.reset();
    h->drop_prefetched();
    h->spare.reset();
    return 0;
}
int cm31_input_prefetch(cm31_prover_input* h) {
    try {
        CM_REQUIRE(h != nullptr, "input_prefetch: null handle");
        CM_REQUIRE(!h->device_adapted, "input_prefetch: a device-adapted input is already resident");
        CM_REQUIRE(h->prefetched.size() < 4, "input_prefetch: too many uploads in flight");
        static const bool slots = getenv("CM31_NO_INPUT_SLOTS") == nullptr;
        std::unique_ptr<StagedInput<CudaAirImpl>> spare = std::move(h->spare);
        // The bulk copies are recorded, not issued: a proof that is running (or about to run) releases them where it is not
        // launch-bound (CudaAirImpl::staging_release_point), the consumer releases them at the latest.
        uint64_t ticket = 0;
        const bool defer = CudaAirImpl::prefetch_point() >= 0;
        if (defer) cm_check(cm31_bg_defer(1, &ticket));
        try {
            h->prefetched.emplace_back(new StagedInput<CudaAirImpl>(stage_input<CudaAirImpl>(h->input, slots ? spare.get() : nullptr)));
        } catch (...) {
            if (defer) cm31_bg_defer(0, nullptr);
            throw;
        }
        h->prefetched.back()->bg_ticket = ticket;
        if (defer) cm_check(cm31_bg_defer(0, nullptr));
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
// info[0] = VM steps, info[1] = data accesses, info[2] = boundary memory rows, info[3] = return value,
// info[4] = bytes of prover input copied host->device per proof
int cm31_input_info(const cm31_prover_input* h, uint64_t info[5]) {
    CM_REQUIRE(h != nullptr, "input_info: null handle");
    info[0] = h->input.n_steps;
    info[1] = h->input.data_accesses.size();
    info[2] = h->input.initial_memory.size() + h->input.final_memory.size();
    info[3] = h->return_value;
    info[4] = h->input.n_steps * sizeof(Bundle) + h->input.data_accesses.size() * sizeof(DataAccess) + info[2] * 32 +
              h->input.clock_update_data.size() * 24;
    if (h->device_adapted) {
        info[1] = h->adapted_info[0];
        info[4] = h->adapted_info[1];
    }
    return 0;
}

// ProverInput from caller-owned flat tables / the flat view of a handle (include/cm31.h: cm31_prover_input_desc)
static_assert(sizeof(Bundle) == 48 && sizeof(DataAccess) == 16 && sizeof(MemoryRow) == 32 && sizeof(ClockUpdateRow) == 24 && sizeof(MerkleNode) == 36,
              "prover-input records must be packed u32 rows");
int cm31_input_create(const cm31_prover_input_desc* d, cm31_prover_input** out) {
    try {
        CM_REQUIRE(d != nullptr && out != nullptr, "input_create: null argument");
        CM_REQUIRE(d->n_steps >= 1, "adapter: empty trace");
        CM_REQUIRE(d->n_opcodes == 0 || (d->opcode_ids && d->bundle_start && d->bundles), "input_create: null bundle tables");
        CM_REQUIRE((d->n_data_accesses == 0 || d->data_accesses) && (d->n_initial_memory == 0 || d->initial_memory) &&
                       (d->n_final_memory == 0 || d->final_memory) && (d->n_clock_updates == 0 || d->clock_updates) &&
                       (d->n_merkle_nodes == 0 || d->merkle_nodes),
                   "input_create: null table with a non-zero count");
        std::unique_ptr<cm31_prover_input> h(new cm31_prover_input());
        ProverInput& in = h->input;
        in.initial_registers = Registers{d->initial_pc, d->initial_fp};
        in.final_registers = Registers{d->final_pc, d->final_fp};
        in.public_ranges.program_start = d->public_ranges[0];
        in.public_ranges.program_end = d->public_ranges[1];
        in.public_ranges.input_start = d->public_ranges[2];
        in.public_ranges.input_end = d->public_ranges[3];
        in.public_ranges.output_start = d->public_ranges[4];
        in.public_ranges.output_end = d->public_ranges[5];
        in.initial_root = d->initial_root;
        in.final_root = d->final_root;
        in.n_steps = d->n_steps;
        uint64_t total = 0;
        for (uint64_t g = 0; g < d->n_opcodes; g++) {
            uint32_t op = d->opcode_ids[g];
            CM_REQUIRE(opcode_memory_accesses(op) >= 0, "adapter: invalid opcode");
            CM_REQUIRE(d->bundle_start[g] <= d->bundle_start[g + 1] && !in.states_by_opcodes.count(op), "input_create: bad opcode groups");
            const Bundle* b = (const Bundle*)d->bundles;
            std::vector<Bundle>& v = in.states_by_opcodes[op];
            v.assign(b + d->bundle_start[g], b + d->bundle_start[g + 1]);
            for (const Bundle& x : v)
                CM_REQUIRE((uint64_t)x.span_start + x.span_len <= d->n_data_accesses && x.span_len <= MAX_ACCESSES, "input_create: access span outside the data-access log");
            total += v.size();
        }
        CM_REQUIRE(total == d->n_steps, "input_create: the opcode groups do not add up to n_steps");
        in.data_accesses.assign((const DataAccess*)d->data_accesses, (const DataAccess*)d->data_accesses + d->n_data_accesses);
        in.initial_memory.assign((const MemoryRow*)d->initial_memory, (const MemoryRow*)d->initial_memory + d->n_initial_memory);
        in.final_memory.assign((const MemoryRow*)d->final_memory, (const MemoryRow*)d->final_memory + d->n_final_memory);
        in.clock_update_data.assign((const ClockUpdateRow*)d->clock_updates, (const ClockUpdateRow*)d->clock_updates + d->n_clock_updates);
        in.merkle_nodes.assign((const MerkleNode*)d->merkle_nodes, (const MerkleNode*)d->merkle_nodes + d->n_merkle_nodes);
        pin_range(h.get(), in.data_accesses.data(), in.data_accesses.size() * sizeof(DataAccess));
        for (auto& kv : in.states_by_opcodes) pin_range(h.get(), kv.second.data(), kv.second.size() * sizeof(Bundle));
        *out = h.release();
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
int cm31_input_describe(cm31_prover_input* h, cm31_prover_input_desc* d) {
    CM_REQUIRE(h != nullptr && d != nullptr, "input_describe: null argument");
    CM_REQUIRE(!h->device_adapted, "input_describe: the per-step tables of a device-adapted input exist in HBM only (cm31_input_staged_words)");
    const ProverInput& in = h->input;
    h->desc_ids.clear();
    h->desc_bundles.clear();
    h->desc_start.assign(1, 0);
    for (const auto& kv : in.states_by_opcodes) {
        h->desc_ids.push_back(kv.first);
        const uint32_t* w = (const uint32_t*)kv.second.data();
        h->desc_bundles.insert(h->desc_bundles.end(), w, w + 12 * kv.second.size());
        h->desc_start.push_back(h->desc_start.back() + kv.second.size());
    }
    memset(d, 0, sizeof(*d));
    d->initial_pc = in.initial_registers.pc;
    d->initial_fp = in.initial_registers.fp;
    d->final_pc = in.final_registers.pc;
    d->final_fp = in.final_registers.fp;
    const PublicRanges& r = in.public_ranges;
    uint32_t ranges[6] = {r.program_start, r.program_end, r.input_start, r.input_end, r.output_start, r.output_end};
    memcpy(d->public_ranges, ranges, sizeof(ranges));
    d->initial_root = in.initial_root;
    d->final_root = in.final_root;
    d->n_steps = in.n_steps;
    d->n_opcodes = h->desc_ids.size();
    d->opcode_ids = h->desc_ids.data();
    d->bundle_start = h->desc_start.data();
    d->bundles = h->desc_bundles.data();
    d->data_accesses = (const uint32_t*)in.data_accesses.data();
    d->n_data_accesses = in.data_accesses.size();
    d->initial_memory = (const uint32_t*)in.initial_memory.data();
    d->n_initial_memory = in.initial_memory.size();
    d->final_memory = (const uint32_t*)in.final_memory.data();
    d->n_final_memory = in.final_memory.size();
    d->clock_updates = (const uint32_t*)in.clock_update_data.data();
    d->n_clock_updates = in.clock_update_data.size();
    d->merkle_nodes = (const uint32_t*)in.merkle_nodes.data();
    d->n_merkle_nodes = in.merkle_nodes.size();
    return 0;
}

// ------------------------------------------------------------------ runner logs + device adapter
}  // extern "C"
struct cm31_test_vm_trace {
    cm31::VmTrace vm;
    std::vector<uint32_t> trace_words;  // IoTraceEntry {fp, pc} per entry (crates/prover/src/adapter/io.rs:38-43)
};
static_assert(sizeof(std::pair<cm31::u32, cm31::Word4>) == 20, "memory-trace entries must be IoMemoryEntry-shaped (5 words)");
extern "C" {

// The runner's output for one of the built-in programs (what cairo-m-runner hands to import_from_runner_output):
// execution trace, memory-access log, preloaded memory, public address ranges.
int cm31_test_vm_trace_create(uint32_t program_id, uint32_t n, cm31_test_vm_trace** out) {
    try {
        CM_REQUIRE(out != nullptr, "vm_trace_create: null out");
        std::unique_ptr<cm31_test_vm_trace> h(new cm31_test_vm_trace());
        h->vm = run_program(program_by_id(program_id), n);
        h->trace_words.reserve(h->vm.trace.size() * 2);
        for (const Registers& r : h->vm.trace) {
            h->trace_words.push_back(r.fp);
            h->trace_words.push_back(r.pc);
        }
        *out = h.release();
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
// info[0] = trace entries (steps + 1), [1] = memory-log entries, [2] = preloaded cells, [3] = return value
int cm31_test_vm_trace_info(const cm31_test_vm_trace* h, uint64_t info[4]) {
    CM_REQUIRE(h != nullptr && info != nullptr, "vm_trace_info: null argument");
    info[0] = h->vm.trace.size();
    info[1] = h->vm.memory_trace.size();
    info[2] = h->vm.initial_memory.size();
    info[3] = h->vm.return_value;
    return 0;
}
int cm31_test_vm_trace_data(const cm31_test_vm_trace* h, const uint32_t** trace, const uint32_t** memory_trace, const uint32_t** initial_memory,
                       uint32_t public_ranges[6]) {
    CM_REQUIRE(h != nullptr, "vm_trace_data: null handle");
    if (trace) *trace = h->trace_words.data();
    if (memory_trace) *memory_trace = (const uint32_t*)h->vm.memory_trace.data();
    if (initial_memory) *initial_memory = (const uint32_t*)h->vm.initial_memory.data();
    if (public_ranges) {
        const PublicRanges& r = h->vm.public_ranges;
        uint32_t v[6] = {r.program_start, r.program_end, r.input_start, r.input_end, r.output_start, r.output_end};
        memcpy(public_ranges, v, sizeof(v));
    }
    return 0;
}
int cm31_test_vm_trace_destroy(cm31_test_vm_trace* h) {
    delete h;
    return 0;
}
// Continuation segments (crates/runner/src/vm/mod.rs:158-285): the run of a built-in program cut every `segment_steps` steps,
// as RunnerOptions::max_steps cuts it; returns segment `index` (its own trace, memory log and initial-memory image) and the
// number of segments.  The reference chains the proofs of consecutive segments by their memory roots
// (crates/prover/tests/prover.rs:204-243).
int cm31_test_vm_segment_create(uint32_t program_id, uint32_t n, uint64_t segment_steps, uint32_t index, uint32_t* n_segments,
                                cm31_test_vm_trace** out) {
    try {
        CM_REQUIRE(out != nullptr && segment_steps > 0, "vm_segment_create: bad arguments");
        std::vector<VmTrace> segs;
        VmTrace last = run_program(program_by_id(program_id), n, (size_t)1 << 30, &segs, (size_t)segment_steps);
        segs.push_back(std::move(last));
        if (n_segments) *n_segments = (uint32_t)segs.size();
        CM_REQUIRE(index < segs.size(), "vm_segment_create: no such segment");
        std::unique_ptr<cm31_test_vm_trace> h(new cm31_test_vm_trace());
        h->vm = std::move(segs[index]);
        h->trace_words.reserve(h->vm.trace.size() * 2);
        for (const Registers& r : h->vm.trace) {
            h->trace_words.push_back(r.fp);
            h->trace_words.push_back(r.pc);
        }
        *out = h.release();
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
// The host adapter (csrc/cairo/vm.hpp::import_from_vm, the restatement of import_from_runner_output) on a given runner output.
int cm31_test_vm_trace_to_input(const cm31_test_vm_trace* t, cm31_prover_input** out) {
    try {
        CM_REQUIRE(t != nullptr && out != nullptr, "vm_trace_to_input: null argument");
        cm31_prover_input* h = new cm31_prover_input();
        h->input = import_from_vm(t->vm);
        h->return_value = t->vm.return_value;
        pin_range(h, h->input.data_accesses.data(), h->input.data_accesses.size() * sizeof(DataAccess));
        for (auto& kv : h->input.states_by_opcodes) pin_range(h, kv.second.data(), kv.second.size() * sizeof(Bundle));
        *out = h;
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}

// Reads one table of the input resident in HBM back to the host (parity tests compare the host adapter's upload with the
// device adapter's output word for word).  table 0 = data-access log, 1..26 = opcode components (CM31_OPCODE_EVALS order),
// 100 = memory rows, 101 = merkle rows, 102 = clock-update rows, 103 = poseidon2 states.  n_words_out = words of REAL rows.
int cm31_input_staged_words(const cm31_prover_input* h, uint32_t table, uint32_t* out, size_t cap_words, size_t* n_words_out) {
    CM_REQUIRE(h != nullptr && n_words_out != nullptr, "input_staged_words: null argument");
    CM_REQUIRE(h->staged != nullptr, "input_staged_words: the input is not resident (cm31_input_upload first)");
    const StagedInput<CudaAirImpl>& st = *h->staged;
    const DeviceCol* words = nullptr;
    size_t n = 0;
    if (table == 0) words = &st.accesses, n = st.n_accesses * 4;
    else if (table >= 1 && table <= st.opcode.size()) words = &st.opcode[table - 1].words, n = st.opcode[table - 1].n_real * 12;
    else if (table == 100) words = &st.memory.words, n = st.memory.n_real * 8;
    else if (table == 101) words = &st.merkle.words, n = st.merkle.n_real * 9;
    else if (table == 102) words = &st.clock_update.words, n = st.clock_update.n_real * 6;
    else if (table == 103) words = &st.poseidon2.words, n = st.poseidon2.n_real * POSEIDON2_T;
    CM_REQUIRE(words != nullptr, "input_staged_words: no such table");
    *n_words_out = n;
    if (out == nullptr || n == 0) return 0;
    CM_REQUIRE(n <= cap_words, "input_staged_words: buffer too small");
    if (int e = cm31_bg_fence()) return e;
    if (int e = cm31_d2h(out, words->ptr(), n * 4)) return e;
    return cm31_sync();
}

// import_from_runner_output (crates/prover/src/adapter/mod.rs:233-…) with the per-step work on the device (csrc/adapter.cu):
// the returned handle is resident in HBM and proves with cm31_prove_cairo_m like an uploaded one.
}  // extern "C"
struct cm31_adapter_logs {  // cm31_adapter_prefetch: the logs on their way to HBM + what the host tail needs
    void* plan = nullptr;
    const uint32_t *trace = nullptr, *initial_memory = nullptr;
    size_t n_trace = 0, n_mem = 0, n_initial = 0;
    uint32_t ranges[6] = {0, 0, 0, 0, 0, 0};
    uint64_t bg_ticket = 0;  // deferred upload (cm31_bg_defer): released by a running proof or by cm31_adapter_import_prefetched
    ~cm31_adapter_logs() {
        if (plan) cm31_adapter_free(plan);
    }
};
// phases 1-2 + the host tail, from logs already staged (consumes logs->plan)
static int adapter_finish(cm31_adapter_logs& logs, cm31_prover_input** out) {
    try {
        uint64_t counts[67];
        if (int e = cm31_adapter_scan_staged(logs.plan, counts)) return e;
        const uint32_t* trace = logs.trace;
        const size_t n_trace = logs.n_trace, n_initial = logs.n_initial;
        std::unique_ptr<cm31_prover_input> h(new cm31_prover_input());
        h->device_adapted = true;
        ProverInput& in = h->input;
        in.initial_registers = Registers{trace[1], trace[0]};
        in.final_registers = Registers{trace[2 * (n_trace - 1) + 1], trace[2 * (n_trace - 1)]};
        in.n_steps = n_trace - 1;
        typedef StagedInput<CudaAirImpl> Staged;
        std::unique_ptr<Staged> st(new Staged());
        st->n_accesses = counts[64];
        st->accesses = CudaAirImpl::alloc_words(st->n_accesses * 4);
        uint32_t* rows_by_opcode[64] = {nullptr};
        auto opcode_rows = [&](const std::vector<u32>& opcodes) {  // one buffer per component, its opcodes' groups in list order
            Staged::Rows r;
            for (u32 op : opcodes) r.n_real += counts[op];
            r.words = CudaAirImpl::alloc_words(r.n_real * 12);
            size_t at = 0;
            for (u32 op : opcodes) {
                if (counts[op]) rows_by_opcode[op] = r.words.ptr() + 12 * at;
                at += counts[op];
            }
            st->opcode.push_back(std::move(r));
        };
#define CM31_X(E) opcode_rows(E::opcodes());
        CM31_OPCODE_EVALS(CM31_X)
#undef CM31_X
        st->clock_update.n_real = counts[65];
        st->clock_update.words = CudaAirImpl::alloc_words(counts[65] * 6);
        std::vector<uint32_t> cells(10 * counts[66] + 1);
        int rc = cm31_adapter_emit(logs.plan, rows_by_opcode, counts, st->accesses.ptr(), st->clock_update.words.ptr(), cells.data());
        cm31_adapter_free(logs.plan);
        logs.plan = nullptr;
        if (rc) return rc;
        // boundary memory from the distinct cells (Memory::push's bookkeeping of initial_memory / final_memory)
        std::vector<Word4> preloaded(n_initial);
        memcpy(preloaded.data(), logs.initial_memory, n_initial * 16);
        MemoryModel memory(preloaded);
        // the memory commitment is a depth-30 tree over 4 words per cell: addresses live below 2^28 (adapter/merkle.rs
        // TREE_HEIGHT).  The cells come back in ascending address order, so the last one is the largest: a stray address in
        // the runner's log is refused here instead of sizing the dense cell tables by it (2 x 2^31 x 28 bytes).
        if (counts[66] != 0 && cells[10 * (counts[66] - 1)] >= (1u << 28)) {
            set_error("cm31: adapter: memory address outside the 2^28-cell address space (VmImportError)");
            return -1;
        }
        for (size_t c = 0; c < counts[66]; c++) {
            const uint32_t* w = &cells[10 * c];
            uint32_t a = w[0];
            if (a >= memory.final_.size()) {
                memory.final_.resize((size_t)a + 1);
                memory.initial.resize((size_t)a + 1);
            }
            if (memory.initial[a].present) memory.initial[a].multiplicity = 1;
            else {
                memory.initial[a].value = Word4{{w[1], w[2], w[3], w[4]}};
                memory.initial[a].clock = 0;
                memory.initial[a].multiplicity = 1;
                memory.initial[a].present = true;
            }
            memory.final_[a].value = Word4{{w[5], w[6], w[7], w[8]}};
            memory.final_[a].clock = w[9];
            memory.final_[a].multiplicity = P - 1;
            memory.final_[a].present = true;
        }
        PublicRanges ranges;
        ranges.program_start = logs.ranges[0];
        ranges.program_end = logs.ranges[1];
        ranges.input_start = logs.ranges[2];
        ranges.input_end = logs.ranges[3];
        ranges.output_start = logs.ranges[4];
        ranges.output_end = logs.ranges[5];
        finish_memory(memory, ranges, in);
        stage_boundary_rows<CudaAirImpl>(in, *st);
        h->adapted_info[0] = counts[64];
        h->adapted_info[1] = n_trace * 8 + logs.n_mem * 20 + n_initial * 16;
        h->staged = std::move(st);
        if (int e = cm31_sync()) return e;
        *out = h.release();
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
static int adapter_stage(const uint32_t* trace, size_t n_trace, const uint32_t* memory_trace, size_t n_mem, const uint32_t* initial_memory,
                         size_t n_initial, const uint32_t public_ranges[6], int background, cm31_adapter_logs& logs) {
    CM_REQUIRE(trace && memory_trace && initial_memory && public_ranges, "adapter_import: null argument");
    if (int e = cm31_adapter_stage_logs(trace, n_trace, memory_trace, n_mem, initial_memory, n_initial, background, &logs.plan)) return e;
    logs.trace = trace;
    logs.initial_memory = initial_memory;
    logs.n_trace = n_trace;
    logs.n_mem = n_mem;
    logs.n_initial = n_initial;
    memcpy(logs.ranges, public_ranges, sizeof(logs.ranges));
    return 0;
}
extern "C" {
int cm31_adapter_import(const uint32_t* trace, size_t n_trace, const uint32_t* memory_trace, size_t n_mem, const uint32_t* initial_memory,
                        size_t n_initial, const uint32_t public_ranges[6], cm31_prover_input** out) {
    CM_REQUIRE(out != nullptr, "adapter_import: null argument");
    cm31_adapter_logs logs;
    if (int e = adapter_stage(trace, n_trace, memory_trace, n_mem, initial_memory, n_initial, public_ranges, 0, logs)) return e;
    return adapter_finish(logs, out);
}
// Pipelined form for a prover fed with continuation segments: cm31_adapter_prefetch starts the upload of a segment's logs on
// the background copy stream and returns at once (the logs must stay valid, and should be page-locked, until
// cm31_adapter_import_prefetched returns); the adapter kernels and the proof of the PREVIOUS segment run meanwhile.
int cm31_adapter_prefetch(const uint32_t* trace, size_t n_trace, const uint32_t* memory_trace, size_t n_mem, const uint32_t* initial_memory,
                          size_t n_initial, const uint32_t public_ranges[6], cm31_adapter_logs** out) {
    CM_REQUIRE(out != nullptr, "adapter_prefetch: null argument");
    std::unique_ptr<cm31_adapter_logs> logs(new cm31_adapter_logs());
    // like cm31_input_prefetch: the bulk copies are recorded; a running proof releases them (throttled) at its release point,
    // cm31_adapter_import_prefetched at the latest
    uint64_t ticket = 0;
    const bool defer = CudaAirImpl::prefetch_point() >= 0;
    if (defer) cm31_bg_defer(1, &ticket);
    int e = adapter_stage(trace, n_trace, memory_trace, n_mem, initial_memory, n_initial, public_ranges, 1, *logs);
    if (defer) cm31_bg_defer(0, nullptr);
    if (e) {
        if (ticket) cm31_bg_cancel(ticket);
        return e;
    }
    logs->bg_ticket = ticket;
    *out = logs.release();
    return 0;
}
// consumes (and frees) `logs`, also on error
int cm31_adapter_import_prefetched(cm31_adapter_logs* logs, cm31_prover_input** out) {
    CM_REQUIRE(logs != nullptr && out != nullptr, "adapter_import_prefetched: null argument");
    std::unique_ptr<cm31_adapter_logs> own(logs);
    CM_REQUIRE(own->plan != nullptr, "adapter_import_prefetched: the logs were already consumed");
    if (own->bg_ticket) {
        if (int e = cm31_bg_release(own->bg_ticket)) return e;  // (no-op when a proof released them already)
        own->bg_ticket = 0;
    }
    return adapter_finish(*own, out);
}
int cm31_adapter_logs_destroy(cm31_adapter_logs* logs) {
    if (logs && logs->bg_ticket) cm31_bg_cancel(logs->bg_ticket);  // recorded, never issued
    if (logs && logs->plan) cm31_bg_fence();  // the upload may still be in flight: order the (stream-ordered) frees after it
    delete logs;
    return 0;
}

// prove_cairo_m::<Blake2sMerkleChannel> (crates/prover/src/prover.rs:23) on the CUDA backend.
// timings_ms (optional, 5 doubles): preprocessed, trace, interaction, stark, total.
// AIR shapes of the 34 components as captured from their `evaluate` bodies (the InfoEvaluator pass of
// FrameworkComponent::new, S/constraint_framework/src/component.rs:139-180, info.rs): host-only, no device work.
extern "C++" {
template <class E>
static auto eval_opcodes(const E&, int) -> decltype(E::opcodes()) { return E::opcodes(); }
template <class E>
static std::vector<cm31::u32> eval_opcodes(const E&, long) { return {}; }
}
int cm31_air_shapes(char* buf, size_t cap, size_t* len) {
    try {
        static const char* rel_names[N_CAIRO_RELATIONS] = {"registers", "memory", "merkle", "poseidon2", "range_check_8", "range_check_16",
                                                           "range_check_20", "bitwise"};
        RelationSet dummy;
        for (int r = 0; r < N_CAIRO_RELATIONS; r++) dummy.relations.push_back(RelationElements::dummy(cairo_relation_size(r)));
        std::vector<std::string> names = cairo_component_names();
        std::vector<u32> ls(names.size(), LOG_N_LANES);
        CairoComponents<CudaAirImpl> comps(ls, &dummy);
        std::string out = "{\"relations\": {";
        for (int r = 0; r < N_CAIRO_RELATIONS; r++)
            out += std::string(r ? ", " : "") + "\"" + rel_names[r] + "\": " + std::to_string(cairo_relation_size(r));
        out += "}, \"components\": [";
        size_t ci = 0;
        comps.for_each([&](auto& c) {
            const ExprEvaluator& ev = c.ev;
            if (ci) out += ", ";
            out += "{\"name\": \"" + names[ci] + "\", \"opcodes\": [";
            std::vector<u32> ops = eval_opcodes(c.eval, 0);
            for (size_t i = 0; i < ops.size(); i++) out += (i ? ", " : "") + std::to_string(ops[i]);
            out += "], \"n_trace_columns\": " + std::to_string(c.n_trace_columns());
            out += ", \"n_interaction_columns\": " + std::to_string(c.n_interaction_columns());
            out += ", \"n_preprocessed_columns\": " + std::to_string(ev.preprocessed_ids.size());
            out += ", \"n_constraints\": " + std::to_string(ev.n_constraints());
            size_t prev_masks = 0;  // interaction columns read at offsets [-1, 0]: the last logup batch (lib.rs:210-234)
            for (auto& offs : ev.mask_offsets[INTERACTION_TRACE_IDX]) prev_masks += offs.size() == 2;
            out += ", \"n_cumsum_columns\": " + std::to_string(prev_masks);
            out += ", \"n_lookups\": " + std::to_string(ev.logup_uses.size()) + ", \"lookups\": {";
            bool first = true;
            for (int r = 0; r < N_CAIRO_RELATIONS; r++) {
                size_t k = 0, widest = 0;
                for (auto& u : ev.logup_uses)
                    if (u.relation == r) k++, widest = std::max(widest, u.values.size());
                if (!k) continue;
                out += std::string(first ? "" : ", ") + "\"" + rel_names[r] + "\": [" + std::to_string(k) + ", " + std::to_string(widest) + "]";
                first = false;
            }
            out += "}}";
            ci++;
        });
        out += "]}";
        if (len) *len = out.size();
        CM_REQUIRE(buf == nullptr || out.size() < cap, "air_shapes: buffer too small");
        if (buf) memcpy(buf, out.c_str(), out.size() + 1);
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}

// ---- asynchronous proofs: the tail of proof i (assembly of the decommitments from the gathered words, serialisation: host
// work with the GPU idle, ~1.4 ms of a 22 ms proof) runs while proof i+1 executes.  One proof may be pending at a time.
struct PendingProofOut {
    std::unique_ptr<CairoProof> proof;
    uint8_t* out = nullptr;
    size_t cap = 0;
    size_t* len = nullptr;
};
static PendingProofOut g_pending_out;
static int g_tail_rc = 0;  // status of the last deferred tail that ran; reported (and cleared) by take_tail_status()
static std::string g_tail_err;
static int take_tail_status() {
    int rc = g_tail_rc;
    if (rc) set_error("deferred tail of an asynchronous proof: " + g_tail_err);
    g_tail_rc = 0;
    g_tail_err.clear();
    return rc;
}
static bool finish_pending_stage() {  // one stage of the pending proof's tail; true once nothing is pending
    PendingProofOut& p = g_pending_out;
    if (!p.proof) return true;
    try {
        if (!p.proof->stark_proof.step()) return false;
        HostTimer ht("proof_to_bytes");
        static thread_local ProofWriter writer;
        writer.bytes.clear();
        p.proof->write(writer);
        int rc = write_out(writer.bytes, p.out, p.cap, p.len);
        if (rc) {
            g_tail_rc = rc;
            g_tail_err = cm31_last_error();
        }
    } catch (const std::exception& e) {
        g_tail_rc = -2;
        g_tail_err = e.what();
    }
    p.proof.reset();
    return true;
}
static void finish_pending_proof() {
    while (!finish_pending_stage()) {
    }
}
static int prove_impl(const cm31_prover_input* h, uint32_t pow_bits, uint32_t n_queries, uint8_t* proof_out, size_t proof_cap,
                      size_t* proof_len, double* timings_ms, bool async);
// Returns once the proof-of-work nonce of the proof is known; proof_out / proof_len are written later -- during the next
// cm31_prove_cairo_m[_async] call on this thread or by cm31_prove_wait(), whichever comes first -- and must stay valid until
// then.  When call i+1 returns 0, proof i is complete and its bytes are valid; a failure of a deferred part is reported by
// the call that ran it ("deferred tail of an asynchronous proof: ..").
int cm31_prove_cairo_m_async(const cm31_prover_input* h, uint32_t pow_bits, uint32_t n_queries, uint8_t* proof_out,
                             size_t proof_cap, size_t* proof_len, double* timings_ms) {
    return prove_impl(h, pow_bits, n_queries, proof_out, proof_cap, proof_len, timings_ms, true);
}
int cm31_prove_wait(void) {
    CudaBackend::finish_deferred_tails();
    finish_pending_proof();  // (no hook registered: e.g. the proof failed before its tail was deferred)
    return take_tail_status();
}
int cm31_prove_cairo_m(const cm31_prover_input* h, uint32_t pow_bits, uint32_t n_queries, uint8_t* proof_out,
                       size_t proof_cap, size_t* proof_len, double* timings_ms) {
    if (int e = cm31_prove_wait()) return e;  // a pending asynchronous proof is completed first
    return prove_impl(h, pow_bits, n_queries, proof_out, proof_cap, proof_len, timings_ms, false);
}
static int prove_impl(const cm31_prover_input* h, uint32_t pow_bits, uint32_t n_queries, uint8_t* proof_out, size_t proof_cap,
                      size_t* proof_len, double* timings_ms, bool async) {
    struct DeferGuard {  // prove_values defers its tail only inside an asynchronous proof
        ~DeferGuard() { CudaBackend::tail_state().defer = false; }
    } defer_guard;
    try {
        if (int rc = take_tail_status()) return rc;  // a deferred tail failed since the last call: report it now
        {  // pool head-room (cm31_pool_reserve_headroom) from the high-water mark of the proofs made so far: before the second
           // and the third proof of the process -- i.e. inside any warm-up -- and again only if a later proof at least doubles
           // the mark (a much larger input)
            static uint64_t n_proofs = 0, seen_high = 0;
            static const bool off = getenv("CM31_NO_POOL_HEADROOM") != nullptr;
            if (!off && n_proofs > 0) {
                int dev = 0;
                cudaMemPool_t pool;
                uint64_t high = 0;
                if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess &&
                    cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &high) == cudaSuccess) {
                    if (n_proofs <= 2 || high >= 2 * seen_high) {
                        seen_high = std::max(seen_high, high);
                        cm31_pool_reserve_headroom(n_proofs == 1 ? 2.5 : 1.5);
                    }
                } else {
                    cudaGetLastError();
                }
            }
            n_proofs++;
        }
        CudaBackend::tail_state().defer = async;
        CM_REQUIRE(h != nullptr, "prove_cairo_m: null input");
        PcsConfig cfg = PcsConfig::regular_96_bits();
        cfg.pow_bits = pow_bits;
        cfg.fri_config.n_queries = n_queries;
        ProveTimings t;
        std::unique_ptr<StagedInput<CudaAirImpl>> pre;
        if (!h->staged && !h->prefetched.empty()) {
            pre = std::move(const_cast<cm31_prover_input*>(h)->prefetched.front());
            const_cast<cm31_prover_input*>(h)->prefetched.pop_front();
            if (pre->bg_ticket) cm_check(cm31_bg_release(pre->bg_ticket));  // its own copies, if no earlier proof released them
        }
        CairoProof proof = h->staged ? prove_cairo_m<CudaAirImpl>(h->input, *h->staged, cfg, &t)
                           : pre     ? prove_cairo_m<CudaAirImpl>(h->input, *pre, cfg, &t)
                                     : prove_cairo_m<CudaAirImpl>(h->input, cfg, &t);
        if (timings_ms) {
            timings_ms[0] = t.preprocessed_ms;
            timings_ms[1] = t.trace_ms;
            timings_ms[2] = t.interaction_ms;
            timings_ms[3] = t.stark_ms;
            timings_ms[4] = t.total_ms;
        }
        if (pre) const_cast<cm31_prover_input*>(h)->spare = std::move(pre);  // keep the slot for the next prefetch
        if (async && proof.stark_proof.pending_tail) {
            // (the previous pending proof was completed inside this proof at the latest: prove_values runs
            // finish_deferred_tails before it takes over the gather landing buffer)
            finish_pending_proof();
            g_pending_out.proof.reset(new CairoProof(std::move(proof)));
            g_pending_out.out = proof_out;
            g_pending_out.cap = proof_cap;
            g_pending_out.len = proof_len;
            CudaBackend::tail_state().hook = finish_pending_stage;
            HostTimer::report();
            // the previous asynchronous proof was completed inside this call: a zero return also vouches for ITS bytes
            return take_tail_status();
        }
        proof.stark_proof.resolve();
        int rc;
        {
            HostTimer ht("proof_to_bytes");
            static thread_local ProofWriter writer;  // reused across proofs: a fresh megabyte costs ~250 page faults per proof
            writer.bytes.clear();
            proof.write(writer);
            rc = write_out(writer.bytes, proof_out, proof_cap, proof_len);
        }
        HostTimer::report();
        return rc;
    } catch (const std::exception& e) {
        {  // programs recorded for a batch that will never be issued point into columns that are gone
            CudaBackend::AirBatch& b = CudaBackend::air_batch();
            b.items.clear();
            b.finals.clear();
            b.depth = 0;
        }
        cm31_lanes_join();  // an error may have been raised while the side lane was current
        set_error(e.what());
        return -2;
    }
}

// Host-side plan of a sharded proof (no device work): which rank owns each of the 34 components for given per-component
// costs, and the node range of a striped layer -- what every rank computes for itself; exported so the multi-process tests
// can check that all ranks agree.
int cm31_shard_plan(const double* cost, size_t n_components, int world, int* owners_out) {
    CM_REQUIRE(cost != nullptr && owners_out != nullptr && world >= 1, "shard_plan: bad arguments");
    std::vector<int> owners = assign_component_owners(std::vector<double>(cost, cost + n_components), world);
    for (size_t i = 0; i < n_components; i++) owners_out[i] = owners[i];
    return 0;
}

// ---- the reference's wire format (serde JSON of Proof<Blake2sMerkleHasher>, crates/prover/src/lib.rs:61-73)
int cm31_proof_to_json(const uint8_t* proof, size_t proof_len, char* json_out, size_t cap, size_t* json_len) {
    try {
        CM_REQUIRE(proof != nullptr, "proof_to_json: null proof");
        std::string js = cairo_proof_to_json(CairoProof::from_bytes(proof, proof_len, cairo_component_names()));
        if (json_len) *json_len = js.size();
        CM_REQUIRE(json_out == nullptr || js.size() < cap, "proof_to_json: buffer too small");
        if (json_out) memcpy(json_out, js.c_str(), js.size() + 1);
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
int cm31_proof_from_json(const char* json, size_t json_len, uint8_t* proof_out, size_t cap, size_t* proof_len) {
    try {
        CM_REQUIRE(json != nullptr, "proof_from_json: null text");
        return write_out(cairo_proof_from_json(json, json_len).to_bytes(), proof_out, cap, proof_len);
    } catch (const std::exception& e) {
        set_error(e.what());
        return -2;
    }
}
int cm31_prove_cairo_m_json(const cm31_prover_input* h, uint32_t pow_bits, uint32_t n_queries, char* json_out, size_t cap, size_t* json_len) {
    // proofs are deterministic: a NULL json_out proves once to learn the length, like the other size queries of this ABI;
    // callers that care pass a buffer of 8x the blob size (the JSON is ~3.4x the blob)
    std::vector<uint8_t> blob((size_t)1 << 24);
    size_t n = 0;
    int rc = cm31_prove_cairo_m(h, pow_bits, n_queries, blob.data(), blob.size(), &n, nullptr);
    if (rc == -1 && n > blob.size()) {
        blob.resize(n);
        rc = cm31_prove_cairo_m(h, pow_bits, n_queries, blob.data(), blob.size(), &n, nullptr);
    }
    if (rc) return rc;
    return cm31_proof_to_json(blob.data(), n, json_out, cap, json_len);
}

}  // extern "C"
