// Build-time generator of the AOT-specialised AIR kernels (cairo-m_b200/csrc/generated/*.cu).
//
// For every cairo-m component it captures the AIR exactly like the prover does at run time
// (FrameworkEval::evaluate -> ExprEvaluator graph, external/stwo/crates/constraint_framework/src/
// expr/evaluator.rs:63-260 is the reference's symbolic evaluator) and writes each of the
// component's programs — trace fill, logup columns, lookup histograms, constraint evaluation — as
// a straight-line CUDA kernel keyed by the hash of the bytecode it is equivalent to.
//   g++ -std=c++17 -I cairo-m_b200/csrc tools/gen_air_kernels.cpp -o build/gen_air_kernels && build/gen_air_kernels <outdir>
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <map>
#include <set>
#include <sstream>

#include "air/cairo_components.hpp"
#include "host/framework.hpp"

using namespace cm31;

struct Kernel {
    std::string name;
    AirProgram prog;
    size_t n_in = 0, n_out = 0, n_consts = 0;  // what the instruction words actually reference
    bool constraint = false;
};

static void shape_of(Kernel& k) {
    for (uint64_t ins : k.prog.code) {
        u32 op = (u32)(ins & 0xff), a = (u32)((ins >> 24) & 0xfffff), b = (u32)((ins >> 44) & 0xfffff);
        if (op == OP_LOAD) k.n_in = std::max<size_t>(k.n_in, a + 1);
        if (op == OP_STORE_E) k.n_out = std::max<size_t>(k.n_out, b + 4);
        if (op == OP_STORE_F || op == OP_HIST) k.n_out = std::max<size_t>(k.n_out, b + 1);
        if (op == OP_CONSTF || op == OP_ROWLT) k.n_consts = std::max<size_t>(k.n_consts, a + 1);
        if (op == OP_CONSTE) k.n_consts = std::max<size_t>(k.n_consts, a + 4);
        if (op == OP_CONSTRAINT_E || op == OP_CONSTRAINT_F) k.n_consts = std::max<size_t>(k.n_consts, b + 4);
    }
}

static std::set<uint64_t> g_seen;
static std::map<uint64_t, std::vector<uint64_t>> g_code_of;  // collision check
#ifndef GEN_BLOCK
#define GEN_BLOCK 256
#endif
#ifndef GEN_SYNC_EVERY
#define GEN_SYNC_EVERY 48
#endif

static void emit_kernel(std::ostream& os, const Kernel& k, std::vector<std::pair<uint64_t, std::string>>& table) {
    uint64_t h = air_code_hash(k.prog.code.data(), k.prog.code.size());
    if (!g_seen.insert(h).second) {
        // the same hash must mean the same program (one kernel serves both); a collision between different programs would
        // make the second one run the first one's kernel
        if (g_code_of[h] != k.prog.code) {
            fprintf(stderr, "gen_air_kernels: air_code_hash collision between two different programs (%s)\n", k.name.c_str());
            exit(1);
        }
        return;
    }
    g_code_of[h] = k.prog.code;
    const std::string& n = k.name;
    size_t nin = std::max<size_t>(1, k.n_in), nout = std::max<size_t>(1, k.n_out), nc = std::max<size_t>(1, k.n_consts);
    os << "// ---------------------------------------------------------------- " << n << "\n";
    os << "// " << k.prog.code.size() << " bytecode instructions, " << k.prog.n_mul_m31 << " M31 multiplications per row, " << k.n_in
       << " input / " << k.n_out << " output columns\n";
    os << "struct A_" << n << " {\n    u32 row_log, trace_log;\n    const u32* denom_inv;\n    u32* acc[4];\n    const u32* in[" << nin
       << "];\n    u32* out[" << nout << "];\n    u32 c[" << nc << "];\n    u32 hist_bins;\n    u32* err;\n    int sync;\n};\n";
    // CTA size: the programs are thousands of straight-line instructions executed once per row, so instruction fetch is
    // shared only between warps that run the same code at the same time; the warps of one CTA start together and stay
    // close, warps of different CTAs do not (ncu r01b: no_inst 30 % with one warp per scheduler and CTA).
    const unsigned block = GEN_BLOCK;
    os << "__global__ void __launch_bounds__(" << block << ") k_" << n << "(const __grid_constant__ A_" << n << " a) {\n";
    os << "    const u32 row_raw = blockIdx.x * " << block << "u + threadIdx.x;\n    const bool live = row_raw < (1u << a.row_log);\n"
       << "    const u32 row = live ? row_raw : 0u;\n";
    if (k.constraint) os << "    QM31 acc = qm_zero();\n    u64 ca0 = 0, ca1 = 0, ca2 = 0, ca3 = 0;\n    u32 ca_n = 0;\n";
    {  // the body, with a (run-time switchable) CTA barrier every GEN_SYNC_EVERY statements
        std::istringstream body(k.prog.cuda_body);
        std::string line;
        size_t n_lines = 0;
        while (std::getline(body, line)) {
            os << line << "\n";
            if (++n_lines % GEN_SYNC_EVERY == 0) os << "    GEN_SYNC();\n";
        }
    }
    if (k.constraint) {
        // component.rs:413-421: col[row] += row_res * denom_inv[row >> trace_log]
        os << "    acc = qm_add(acc, dot_done(ca0, ca1, ca2, ca3));\n";
        os << "    const QM31 res = qm_mul_m31(acc, __ldg(a.denom_inv + (row >> a.trace_log)));\n";
        os << "    if (live) {\n";
        os << "        a.acc[0][row] = m31_add(a.acc[0][row], res.a);\n        a.acc[1][row] = m31_add(a.acc[1][row], res.b);\n";
        os << "        a.acc[2][row] = m31_add(a.acc[2][row], res.c);\n        a.acc[3][row] = m31_add(a.acc[3][row], res.d);\n";
        os << "    }\n";
    }
    os << "}\n";
    os << "static int l_" << n << "(const GenLaunch& g) {\n";
    os << "    CM_REQUIRE(g.n_in >= " << k.n_in << " && g.n_out >= " << k.n_out << " && g.n_consts >= " << k.n_consts
       << ", \"generated AIR kernel " << n << ": argument shape mismatch\");\n";
    os << "    CM_REQUIRE(" << (k.constraint ? "g.acc4 != nullptr && g.denom_inv_dev != nullptr" : "g.acc4 == nullptr")
       << ", \"generated AIR kernel " << n << ": accumulator mismatch\");\n";
    os << "    A_" << n << " a;\n    a.row_log = g.row_log;\n    a.trace_log = g.trace_log;\n    a.denom_inv = g.denom_inv_dev;\n";
    os << "    a.hist_bins = g.hist_bins;\n    a.err = g.err_flag;\n    a.sync = g.sync;\n";
    os << "    for (int k = 0; k < 4; k++) a.acc[k] = g.acc4 ? g.acc4[k] : nullptr;\n";
    os << "    for (size_t k = 0; k < " << k.n_in << "; k++) a.in[k] = g.in_cols[k];\n";
    os << "    for (size_t k = 0; k < " << k.n_out << "; k++) a.out[k] = g.out_cols[k];\n";
    os << "    for (size_t k = 0; k < " << k.n_consts << "; k++) a.c[k] = g.consts[k];\n";
    os << "    const size_t n = (size_t)1 << g.row_log;\n";
    os << "    k_" << n << "<<<(unsigned)((n + " << (block - 1) << ") / " << block << "), " << block << ", 0, stream()>>>(a);\n    return 0;\n}\n\n";
    char buf[32];
    snprintf(buf, sizeof buf, "0x%016llxull", (unsigned long long)h);
    table.push_back({h, std::string("{") + buf + ", \"" + n + "\", l_" + n + "}"});
}

template <class Eval>
static void component(const std::string& outdir, Eval eval, std::vector<std::string>& accessors, bool has_trace_program = true) {
    std::string cname = Eval::name();
    std::ostringstream os;
    os << "// GENERATED by tools/gen_air_kernels.cpp from csrc/air/cairo_components.hpp (" << cname << ") — do not edit.\n";
    os << "#include \"../air_gen.cuh\"\n\nnamespace cm31 {\nnamespace {\n\n";
    std::vector<std::pair<uint64_t, std::string>> table;
    ExprEvaluator ev;
    eval.evaluate(ev);
    {
        Kernel k{cname + "_constraints", build_constraint_program(ev, true)};
        k.constraint = true;
        shape_of(k);
        emit_kernel(os, k, table);
    }
    // multiplicity histograms exist for the table relations only (opcodes/mod.rs:83-105)
    for (int rel : {REL_RC8, REL_RC16, REL_RC20, REL_BITWISE}) {
        if (!has_trace_program) break;  // the tables themselves look nothing up
        Kernel k{cname + "_lookups_rel" + std::to_string(rel), build_lookup_program(ev, rel, cairo_table_index_weights(rel), true)};
        if (k.prog.code.empty()) continue;
        shape_of(k);
        emit_kernel(os, k, table);
    }
    {
        ExprEvaluator e2 = ev;
        if (!e2.batch_fracs.empty()) {
            Kernel k{cname + "_logup", build_logup_program(e2, true)};
            shape_of(k);
            emit_kernel(os, k, table);
        }
    }
    if constexpr (!std::is_same<Eval, RangeCheckEval>::value && !std::is_same<Eval, BitwiseEval>::value) {
        if (has_trace_program) {
            TraceProgramBuilder tb(1);
            eval.write_trace(tb);
            Kernel k{cname + "_trace", tb.compile(true)};
            shape_of(k);
            emit_kernel(os, k, table);
        }
    }
    os << "const GenEntry table[] = {\n";
    for (auto& t : table) os << "    " << t.second << ",\n";
    os << "    {0, nullptr, nullptr}};\n\n}  // namespace\n\n";
    os << "const GenEntry* air_gen_table_" << cname << "() { return table; }\n\n}  // namespace cm31\n";
    std::string path = outdir + "/air_" + cname + ".cu";
    // rewrite only on change so incremental builds skip nvcc
    std::ifstream old(path);
    std::stringstream cur;
    cur << old.rdbuf();
    if (cur.str() != os.str()) std::ofstream(path) << os.str();
    accessors.push_back("air_gen_table_" + cname);
    fprintf(stderr, "gen_air_kernels: %s: %zu kernels\n", cname.c_str(), table.size());
}

template <class Eval>
static Eval make(u32 log_size = 4) {
    Eval e;
    e.log_size_ = log_size;
    return e;
}

int main(int argc, char** argv) {
    std::string outdir = argc > 1 ? argv[1] : ".";
    std::vector<std::string> acc;
#define CM31_X(E) component(outdir, make<E>(), acc);
    CM31_OPCODE_EVALS(CM31_X)
#undef CM31_X
    component(outdir, make<MemoryEval>(), acc);
    component(outdir, make<MerkleEval>(), acc);
    component(outdir, make<ClockUpdateEval>(), acc);
    component(outdir, make<Poseidon2Eval>(), acc);
    {
        RangeCheckEval e;
        e.log_size_ = 20;
        e.relation = REL_RC20;
        component(outdir, e, acc, false);
    }
    component(outdir, make<BitwiseEval>(BITWISE_STACKED_LOG_SIZE), acc, false);
    std::ostringstream os;
    os << "// GENERATED by tools/gen_air_kernels.cpp — do not edit.\n#include \"../air_gen.cuh\"\n\nnamespace cm31 {\n\n";
    for (auto& a : acc) os << "const GenEntry* " << a << "();\n";
    os << "\nconst GenEntry* air_gen_lookup(uint64_t hash) {\n    const GenEntry* tables[] = {";
    for (auto& a : acc) os << a << "(), ";
    os << "};\n    for (const GenEntry* t : tables)\n        for (; t->launch; t++)\n            if (t->hash == hash) return t;\n    return nullptr;\n}\n\n}  // namespace cm31\n";
    std::string path = outdir + "/air_registry.cu";
    std::ifstream old(path);
    std::stringstream cur;
    cur << old.rdbuf();
    if (cur.str() != os.str()) std::ofstream(path) << os.str();
    return 0;
}
