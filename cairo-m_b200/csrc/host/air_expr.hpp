// Symbolic capture of an AIR (`FrameworkEval::evaluate`) and its compilation to the register
// bytecode executed by the sm_100a AIR kernel (csrc/air.cu).
//
// Reference counterparts: `EvalAtRow` (external/stwo/crates/constraint_framework/src/lib.rs:43-155),
// the logup finalisation `logup_proxy!` (lib.rs:170-250), `LookupElements::combine`
// (logup.rs:96-111) and the symbolic evaluator `ExprEvaluator`
// (constraint_framework/src/expr/evaluator.rs:63-260) which the Rust shim would use to emit this
// bytecode.  One captured graph yields three programs:
//   * constraint program  : Σ α^k·constraint_k per LDE row   (component.rs:283-424)
//   * logup program       : cumulative logup columns per trace row (logup.rs:123-320)
//   * lookup-emission     : values looked up in a table relation -> multiplicity histogram
//                           (crates/prover/src/preprocessed/range_check/range_check_macro.rs:72-84)
#pragma once
#include <cassert>
#include <cstdint>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "qm_ops.hpp"

namespace cm31 {

constexpr int PREPROCESSED_TRACE_IDX = 0;
constexpr int ORIGINAL_TRACE_IDX = 1;
constexpr int INTERACTION_TRACE_IDX = 2;

// ------------------------------------------------------------------ bytecode
enum AirOp : uint8_t {
    OP_LOAD = 1,   // dst <- in_cols[a][row + offset b (signed, in trace-domain steps)]
    OP_CONSTF,     // dst <- consts[a]
    OP_CONSTE,     // dst[0..4] <- consts[a..a+4]
    OP_ADD, OP_SUB, OP_MUL, OP_NEG,          // M31
    OP_EADD, OP_ESUB, OP_EMUL, OP_ENEG,      // QM31 (4 slots)
    OP_EMULF,      // dst[0..4] <- E(a) * F(b)
    OP_EADDF,      // dst[0..4] <- E(a) + F(b)
    OP_ESUBF,      // dst[0..4] <- E(a) - F(b)
    OP_F2E,        // dst[0..4] <- (F(a),0,0,0)
    OP_MOV,        // dst <- F(a)           (used to assemble combine_ef)
    OP_EINV,       // dst[0..4] <- 1 / E(a)
    OP_CONSTRAINT_E,  // acc += consts[b..b+4] * E(a)
    OP_CONSTRAINT_F,  // acc += consts[b..b+4] * F(a)
    OP_STORE_E,    // out_cols[b..b+4][row] <- E(a)
    OP_STORE_F,    // out_cols[b][row] <- F(a)
    OP_HIST,       // atomicAdd(out_cols[b][F(a)], 1)
    OP_INV,        // dst <- F(a)^-1 (0 -> 0)
    OP_SHR,        // dst <- F(a) >> b   (integer op on the canonical representative)
    OP_AND,        // dst <- F(a) & b
    OP_ROWLT,      // dst <- (row < consts[a]) ? 1 : 0     (Enabler, crates/prover/src/utils/enabler.rs:57-75)
    OP_LE,         // dst <- (F(a) <= F(b)) ? 1 : 0   (integer comparison of canonical representatives; witness only)
    OP_DIVC,       // dst <- F(a) / b   (integer division by the constant b; witness only)
    OP_MODC,       // dst <- F(a) % b
    OP_U32DIVREM,  // dst <- 16-bit part b (0: q_lo, 1: q_hi, 2: r_lo, 3: r_hi) of the euclidean division of the u32
                   //        (F(a) | F(a+1) << 16) by (F(a+2) | F(a+3) << 16); divisor 0 gives (0, 0)   (witness only)
};
inline uint64_t air_encode(AirOp op, u32 dst, u32 a, u32 b) {
    return (uint64_t)op | ((uint64_t)(dst & 0xffff) << 8) | ((uint64_t)(a & 0xfffff) << 24) | ((uint64_t)(b & 0xfffff) << 44);
}

struct AirProgram {
    std::vector<uint64_t> code;
    std::vector<u32> consts;  // u32 words; QM31 constants occupy 4 consecutive words
    u32 n_regs = 0;
    // host-filled parameter slots inside `consts` (each 4 words)
    std::vector<u32> param_slots;  // param id -> word offset in consts
    std::vector<u32> rowlt_slots;  // consts words holding a row bound (Enabler): patched per trace
    size_t n_mul_m31 = 0;          // algorithmic M31 multiplications per row (for ops/s reports)
    // Same program as straight-line CUDA (filled when requested): the body of an AOT-specialised
    // kernel (tools/gen_air_kernels.cpp -> csrc/generated/), keyed by `air_code_hash(code)`.
    std::string cuda_body;
};

// Hash of the instruction words: identifies a program independently of its parameter values (all constants are read from the
// `consts` table at run time).  It is recomputed on EVERY program launch (the C ABI takes the bytecode, not a handle), on the
// thread that issues the launches, so it must be cheap: four independent multiply-xorshift lanes over the 64-bit words
// (~2 cycles per instruction; the byte-wise FNV-1a it replaces cost ~32, 0.25 ms per proof in the launch-bound phases).
inline uint64_t air_code_hash(const uint64_t* code, size_t n_instr) {
    const uint64_t K = 0x9E3779B97F4A7C15ull;
    uint64_t h[4] = {0x243F6A8885A308D3ull, 0x13198A2E03707344ull, 0xA4093822299F31D0ull, 0x082EFA98EC4E6C89ull};
    size_t i = 0;
    for (; i + 4 <= n_instr; i += 4)
        for (int k = 0; k < 4; k++) {
            h[k] = (h[k] ^ code[i + k]) * K;
            h[k] ^= h[k] >> 29;
        }
    for (int k = 0; i < n_instr; i++, k++) {
        h[k] = (h[k] ^ code[i]) * K;
        h[k] ^= h[k] >> 29;
    }
    uint64_t r = (uint64_t)n_instr;
    for (int k = 0; k < 4; k++) {
        r = (r ^ h[k]) * K;
        r ^= r >> 32;
    }
    return r;
}

// ------------------------------------------------------------------ graph
enum class NodeOp : uint8_t { Col, ConstF, ConstE, ParamE, AddF, SubF, MulF, NegF, AddE, SubE, MulE, NegE, MulEF, AddEF, SubEF, F2E, Combine4, InvE, InvF, ShrF, AndF, RowLt, LeF, DivCF, ModCF, U32DivRemF };

struct Node {
    NodeOp op;
    bool ext;
    int a = -1, b = -1, c = -1, d = -1;
    u32 fconst = 0;
    QM31 econst = {0, 0, 0, 0};
    int interaction = 0, col = 0, offset = 0;  // Col
    int param = -1;                            // ParamE
};

struct ColumnRef {
    int interaction, col, offset;
    bool operator<(const ColumnRef& o) const { return std::tie(interaction, col, offset) < std::tie(o.interaction, o.col, o.offset); }
};

class Graph {
   public:
    std::vector<Node> nodes;
    int add(const Node& n) {
        Key k = key_of(n);
        auto it = cse_.find(k);
        if (it != cse_.end()) return it->second;
        nodes.push_back(n);
        int id = (int)nodes.size() - 1;
        cse_[k] = id;
        return id;
    }
    bool is_constf(int id, u32* v = nullptr) const {
        if (nodes[id].op != NodeOp::ConstF) return false;
        if (v) *v = nodes[id].fconst;
        return true;
    }
    int constf(u32 v) {
        Node n{NodeOp::ConstF, false};
        n.fconst = v;
        return add(n);
    }
    int conste(QM31 v) {
        Node n{NodeOp::ConstE, true};
        n.econst = v;
        return add(n);
    }
    int parame(int param) {
        Node n{NodeOp::ParamE, true};
        n.param = param;
        return add(n);
    }
    int col(int interaction, int c, int offset) {
        Node n{NodeOp::Col, false};
        n.interaction = interaction;
        n.col = c;
        n.offset = offset;
        return add(n);
    }
    int bin(NodeOp op, bool ext, int a, int b) {
        Node n{op, ext};
        n.a = a;
        n.b = b;
        return add(n);
    }
    int un(NodeOp op, bool ext, int a) {
        Node n{op, ext};
        n.a = a;
        return add(n);
    }
    // F arithmetic with constant folding and identities
    int addf(int a, int b) {
        u32 x, y;
        bool ca = is_constf(a, &x), cb = is_constf(b, &y);
        if (ca && cb) return constf(m31_add(x, y));
        if (ca && x == 0) return b;
        if (cb && y == 0) return a;
        return bin(NodeOp::AddF, false, a, b);
    }
    int subf(int a, int b) {
        u32 x, y;
        bool ca = is_constf(a, &x), cb = is_constf(b, &y);
        if (ca && cb) return constf(m31_sub(x, y));
        if (cb && y == 0) return a;
        if (a == b) return constf(0);
        return bin(NodeOp::SubF, false, a, b);
    }
    int mulf(int a, int b) {
        u32 x, y;
        bool ca = is_constf(a, &x), cb = is_constf(b, &y);
        if (ca && cb) return constf(m31_mul(x, y));
        if ((ca && x == 0) || (cb && y == 0)) return constf(0);
        if (ca && x == 1) return b;
        if (cb && y == 1) return a;
        return bin(NodeOp::MulF, false, a, b);
    }
    int negf(int a) {
        u32 x;
        if (is_constf(a, &x)) return constf(m31_neg(x));
        return un(NodeOp::NegF, false, a);
    }

   private:
    typedef std::tuple<int, int, int, int, int, u32, u32, u32, u32, int, int, int, int> Key;
    static Key key_of(const Node& n) {
        u32 w0 = (n.op == NodeOp::ConstF || n.op == NodeOp::ShrF || n.op == NodeOp::AndF || n.op == NodeOp::RowLt || n.op == NodeOp::DivCF ||
                  n.op == NodeOp::ModCF || n.op == NodeOp::U32DivRemF)
                     ? n.fconst
                     : n.econst.a;
        return Key((int)n.op, n.a, n.b, n.c, n.d, w0, n.econst.b, n.econst.c, n.econst.d, n.interaction, n.col, n.offset, n.param);
    }
    std::map<Key, int> cse_;
};

// ------------------------------------------------------------------ relations
// `relation!` instances of cairo-m (crates/prover/src/relations.rs:7-44); ids index RelationElements.
struct RelationElements {
    QM31 z = {0, 0, 0, 0};
    QM31 alpha = {0, 0, 0, 0};
    std::vector<QM31> alpha_powers;
    // LookupElements::draw (logup.rs:82-95)
    template <class Channel>
    static RelationElements draw(Channel& ch, size_t n) {
        RelationElements r;
        std::vector<QM31> za = ch.draw_secure_felts(2);
        r.z = za[0];
        r.alpha = za[1];
        QM31 cur = qm_one();
        for (size_t i = 0; i < n; i++) {
            r.alpha_powers.push_back(cur);
            cur = cur * r.alpha;
        }
        return r;
    }
    static RelationElements dummy(size_t n) {  // LookupElements::dummy (logup.rs:113-119)
        RelationElements r;
        r.z = qm_make(1, 2, 3, 4);
        r.alpha = qm_one();
        r.alpha_powers.assign(n, qm_one());
        return r;
    }
};

// ------------------------------------------------------------------ logup mixin (lib.rs:170-250)
// Derived must provide: EF ef_zero(); EF ef_sub(EF,EF); EF ef_add(EF,EF); EF ef_mul(EF,EF);
// EF next_extension_interaction_mask(int interaction, offset) variants; add_constraint_ef(EF);
// EF cumsum_shift().
template <class Derived, class F, class EF>
class LogupMixin {
   public:
    struct Frac {
        EF num, den;
    };
    std::vector<Frac> fracs;
    std::vector<Frac> batch_fracs;  // per logup batch: summed (numerator, denominator)
    bool logup_finalized = true;

    void write_logup_frac(EF num, EF den) {
        if (fracs.empty()) logup_finalized = false;
        fracs.push_back(Frac{num, den});
    }
    void finalize_logup_batched(const std::vector<size_t>& batching) {
        Derived& self = static_cast<Derived&>(*this);
        if (logup_finalized) throw std::logic_error("LogupAtRow was already finalized");
        if (batching.size() != fracs.size()) throw std::logic_error("Batching must be of the same length as the number of entries");
        size_t last_batch = 0;
        for (size_t b : batching) last_batch = b > last_batch ? b : last_batch;
        std::vector<std::vector<Frac>> by_batch(last_batch + 1);
        for (size_t i = 0; i < fracs.size(); i++) by_batch[batching[i]].push_back(fracs[i]);
        for (auto& v : by_batch)
            if (v.empty()) throw std::logic_error("Batching must contain all consecutive batches");
        auto sum_fracs = [&](const std::vector<Frac>& v) {
            // Fraction::sum: zero = 0/1, a/b + c/d = (a*d + b*c)/(b*d)  (lookups/utils.rs)
            Frac acc = v[0];
            for (size_t i = 1; i < v.size(); i++) {
                Frac r;
                r.num = self.ef_add(self.ef_mul(acc.num, v[i].den), self.ef_mul(acc.den, v[i].num));
                r.den = self.ef_mul(acc.den, v[i].den);
                acc = r;
            }
            return acc;
        };
        EF prev_col_cumsum = self.ef_zero();
        for (size_t b = 0; b < last_batch; b++) {
            Frac cur = sum_fracs(by_batch[b]);
            batch_fracs.push_back(cur);
            EF cur_cumsum = self.next_extension_interaction_mask_0(INTERACTION_TRACE_IDX);
            EF diff = self.ef_sub(cur_cumsum, prev_col_cumsum);
            prev_col_cumsum = cur_cumsum;
            self.add_constraint_ef(self.ef_sub(self.ef_mul(diff, cur.den), cur.num));
        }
        Frac frac = sum_fracs(by_batch[last_batch]);
        batch_fracs.push_back(frac);
        EF prev_row_cumsum, cur_cumsum;
        self.next_extension_interaction_mask_m1_0(INTERACTION_TRACE_IDX, prev_row_cumsum, cur_cumsum);
        EF diff = self.ef_sub(self.ef_sub(cur_cumsum, prev_row_cumsum), prev_col_cumsum);
        EF fixed_diff = self.ef_add(diff, self.cumsum_shift());
        self.add_constraint_ef(self.ef_sub(self.ef_mul(fixed_diff, frac.den), frac.num));
        logup_finalized = true;
    }
    void finalize_logup_in_pairs() {
        std::vector<size_t> b;
        for (size_t i = 0; i < fracs.size(); i++) b.push_back(i / 2);
        finalize_logup_batched(b);
    }
    void finalize_logup() {
        std::vector<size_t> b;
        for (size_t i = 0; i < fracs.size(); i++) b.push_back(i);
        finalize_logup_batched(b);
    }
};

// ------------------------------------------------------------------ ExprEvaluator
class ExprEvaluator;
struct FExpr {
    ExprEvaluator* ev = nullptr;
    int id = -1;
};
struct EFExpr {
    ExprEvaluator* ev = nullptr;
    int id = -1;
};

struct LogupUse {
    int relation;
    int multiplicity;         // EF node
    std::vector<int> values;  // F nodes
    int denominator;          // EF node: combine(values)
};

constexpr int PARAM_CUMSUM_SHIFT = 0;
constexpr int PARAM_RELATION_BASE = 1;  // then per relation: z, alpha^0.. ; allocated on demand

class ExprEvaluator : public LogupMixin<ExprEvaluator, FExpr, EFExpr> {
   public:
    typedef FExpr F;
    typedef EFExpr EF;
    Graph g;
    std::vector<int> constraints;       // node ids, F or EF
    std::vector<LogupUse> logup_uses;   // in add_to_relation order
    std::vector<std::vector<std::vector<int>>> mask_offsets;  // [interaction][col] -> offsets
    std::vector<std::string> preprocessed_ids;                // in first-use order
    // parameters: id -> description (filled by the component at proving time)
    struct ParamDesc {
        enum Kind { CumsumShift, RelationZ, RelationAlphaPow } kind;
        int relation;
        int power;
    };
    std::vector<ParamDesc> params;

    ExprEvaluator() {
        mask_offsets.resize(3);
        params.push_back({ParamDesc::CumsumShift, -1, 0});
    }

    // ---- EvalAtRow surface (lib.rs:43-155)
    F next_trace_mask() { return next_interaction_mask(ORIGINAL_TRACE_IDX, 0); }
    F get_preprocessed_column(const std::string& id) {
        // Unlike trace columns, preprocessed columns are addressed by id; the component later maps
        // first-use order to global preprocessed indices (component.rs:146-171).
        preprocessed_ids.push_back(id);
        return next_interaction_mask(PREPROCESSED_TRACE_IDX, 0);
    }
    F next_interaction_mask(int interaction, int offset) {
        std::vector<int> offs = {offset};
        return next_interaction_mask_multi(interaction, offs)[0];
    }
    std::vector<F> next_interaction_mask_multi(int interaction, const std::vector<int>& offsets) {
        if ((int)mask_offsets.size() <= interaction) mask_offsets.resize(interaction + 1);
        int c = (int)mask_offsets[interaction].size();
        mask_offsets[interaction].push_back(offsets);
        std::vector<F> out;
        for (int off : offsets) out.push_back(F{this, g.col(interaction, c, off)});
        return out;
    }
    EF next_extension_interaction_mask_0(int interaction) {
        int ids[4];
        for (int k = 0; k < 4; k++) ids[k] = next_interaction_mask(interaction, 0).id;
        return combine_ef(ids);
    }
    void next_extension_interaction_mask_m1_0(int interaction, EF& prev, EF& cur) {
        int p[4], c[4];
        for (int k = 0; k < 4; k++) {
            std::vector<F> m = next_interaction_mask_multi(interaction, {-1, 0});
            p[k] = m[0].id;
            c[k] = m[1].id;
        }
        prev = combine_ef(p);
        cur = combine_ef(c);
    }
    EF combine_ef(const int ids[4]) {
        Node n{NodeOp::Combine4, true};
        n.a = ids[0];
        n.b = ids[1];
        n.c = ids[2];
        n.d = ids[3];
        return EF{this, g.add(n)};
    }
    void add_constraint(F c) { constraints.push_back(c.id); }
    void add_constraint(EF c) { constraints.push_back(c.id); }
    void add_constraint_ef(EF c) { constraints.push_back(c.id); }
    F add_intermediate(F v) { return v; }
    EF add_extension_intermediate(EF v) { return v; }

    F f_const(u32 v) { return F{this, g.constf(v % P)}; }
    F f_const_i(long long v) { return F{this, g.constf(m31_from_i64(v))}; }
    EF ef(F v) { return EF{this, g.un(NodeOp::F2E, true, v.id)}; }
    EF ef_const(QM31 v) { return EF{this, g.conste(v)}; }
    EF ef_one() { return ef_const(qm_one()); }
    EF ef_zero() { return ef_const(qm_zero()); }
    // Base-field operands embedded with ef() keep their cheap form: E*ef(f) is 4 multiplications,
    // not 16 (exact field identities, so every evaluator computes the same values).
    bool is_f2e(EF a, int* f = nullptr) const {
        if (g.nodes[a.id].op != NodeOp::F2E) return false;
        if (f) *f = g.nodes[a.id].a;
        return true;
    }
    bool is_zero_e(EF a) const { return g.nodes[a.id].op == NodeOp::ConstE && qm_is_zero(g.nodes[a.id].econst); }
    EF ef_add(EF a, EF b) {
        int f;
        if (is_zero_e(a)) return b;
        if (is_zero_e(b)) return a;
        if (is_f2e(b, &f)) return EF{this, g.bin(NodeOp::AddEF, true, a.id, f)};
        if (is_f2e(a, &f)) return EF{this, g.bin(NodeOp::AddEF, true, b.id, f)};
        return EF{this, g.bin(NodeOp::AddE, true, a.id, b.id)};
    }
    EF ef_sub(EF a, EF b) {
        int f;
        if (is_f2e(b, &f)) return EF{this, g.bin(NodeOp::SubEF, true, a.id, f)};
        return EF{this, g.bin(NodeOp::SubE, true, a.id, b.id)};
    }
    EF ef_mul(EF a, EF b) {
        int f;
        if (is_f2e(b, &f)) return EF{this, g.bin(NodeOp::MulEF, true, a.id, f)};
        if (is_f2e(a, &f)) return EF{this, g.bin(NodeOp::MulEF, true, b.id, f)};
        return EF{this, g.bin(NodeOp::MulE, true, a.id, b.id)};
    }
    EF ef_neg(EF a) {
        int f;
        if (is_f2e(a, &f)) return EF{this, g.un(NodeOp::F2E, true, g.negf(f))};
        return EF{this, g.un(NodeOp::NegE, true, a.id)};
    }
    EF ef_mul_f(EF a, F b) { return EF{this, g.bin(NodeOp::MulEF, true, a.id, b.id)}; }
    EF ef_add_f(EF a, F b) { return EF{this, g.bin(NodeOp::AddEF, true, a.id, b.id)}; }
    EF cumsum_shift() { return EF{this, g.parame(PARAM_CUMSUM_SHIFT)}; }
    EF ef_inv(EF a) { return EF{this, g.un(NodeOp::InvE, true, a.id)}; }
    // witness-generation-only helpers (trace-fill programs)
    F f_inv(F a) { return F{this, g.un(NodeOp::InvF, false, a.id)}; }
    F f_shr(F a, u32 k) {
        Node n{NodeOp::ShrF, false};
        n.a = a.id;
        n.fconst = k;
        return F{this, g.add(n)};
    }
    F f_and(F a, u32 mask) {
        Node n{NodeOp::AndF, false};
        n.a = a.id;
        n.fconst = mask;
        return F{this, g.add(n)};
    }
    F f_le(F a, F b) { return F{this, g.bin(NodeOp::LeF, false, a.id, b.id)}; }
    F f_divc(F a, u32 c) {
        Node n{NodeOp::DivCF, false};
        n.a = a.id;
        n.fconst = c;
        return F{this, g.add(n)};
    }
    F f_modc(F a, u32 c) {
        Node n{NodeOp::ModCF, false};
        n.a = a.id;
        n.fconst = c;
        return F{this, g.add(n)};
    }
    // euclidean division of two u32s held as 16-bit limb pairs (u32_store_div_fp_fp.rs:360-400): part 0..3 =
    // q_lo, q_hi, r_lo, r_hi
    F f_u32_divrem(F n_lo, F n_hi, F d_lo, F d_hi, u32 part) {
        Node pack{NodeOp::Combine4, true};
        pack.a = n_lo.id;
        pack.b = n_hi.id;
        pack.c = d_lo.id;
        pack.d = d_hi.id;
        Node n{NodeOp::U32DivRemF, false};
        n.a = g.add(pack);
        n.fconst = part;
        return F{this, g.add(n)};
    }
    F row_lt(u32 bound) {
        Node n{NodeOp::RowLt, false};
        n.fconst = bound;
        return F{this, g.add(n)};
    }
    F input(int col) { return F{this, g.col(3, col, 0)}; }  // interaction 3 = trace-fill inputs

    // add_to_relation (lib.rs:113-121) with LookupElements::combine (logup.rs:96-111)
    void add_to_relation(int relation, EF multiplicity, const std::vector<F>& values) {
        EF acc = ef_zero();
        for (size_t i = 0; i < values.size(); i++) {
            EF power = EF{this, g.parame(relation_param(relation, (int)i + 1))};
            acc = ef_add(acc, ef_mul_f(power, values[i]));
        }
        EF z = EF{this, g.parame(relation_param(relation, 0))};
        EF den = ef_sub(acc, z);
        LogupUse u;
        u.relation = relation;
        u.multiplicity = multiplicity.id;
        for (const F& v : values) u.values.push_back(v.id);
        u.denominator = den.id;
        logup_uses.push_back(u);
        write_logup_frac(multiplicity, den);
    }

    size_t n_constraints() const { return constraints.size(); }

   private:
    // which = 0 -> z, which = i+1 -> alpha^i
    int relation_param(int relation, int which) {
        auto key = std::make_pair(relation, which);
        auto it = rel_params_.find(key);
        if (it != rel_params_.end()) return it->second;
        ParamDesc d;
        d.kind = which == 0 ? ParamDesc::RelationZ : ParamDesc::RelationAlphaPow;
        d.relation = relation;
        d.power = which - 1;
        params.push_back(d);
        int id = (int)params.size() - 1;
        rel_params_[key] = id;
        return id;
    }
    std::map<std::pair<int, int>, int> rel_params_;
};

inline FExpr operator+(FExpr a, FExpr b) { return FExpr{a.ev, a.ev->g.addf(a.id, b.id)}; }
inline FExpr operator-(FExpr a, FExpr b) { return FExpr{a.ev, a.ev->g.subf(a.id, b.id)}; }
inline FExpr operator*(FExpr a, FExpr b) { return FExpr{a.ev, a.ev->g.mulf(a.id, b.id)}; }
inline FExpr operator-(FExpr a) { return FExpr{a.ev, a.ev->g.negf(a.id)}; }
inline EFExpr operator+(EFExpr a, EFExpr b) { return a.ev->ef_add(a, b); }
inline EFExpr operator-(EFExpr a, EFExpr b) { return a.ev->ef_sub(a, b); }
inline EFExpr operator*(EFExpr a, EFExpr b) { return a.ev->ef_mul(a, b); }
inline EFExpr operator-(EFExpr a) { return a.ev->ef_neg(a); }
inline EFExpr operator*(EFExpr a, FExpr b) { return a.ev->ef_mul_f(a, b); }
inline EFExpr operator+(EFExpr a, FExpr b) { return a.ev->ef_add_f(a, b); }

// ------------------------------------------------------------------ host interpretation (QM31)
// Evaluates graph nodes with every value in QM31: this is the PointEvaluator
// (constraint_framework/src/point.rs) used for the OODS sanity check in `prove`.
struct GraphPointEval {
    const Graph& g;
    std::vector<QM31> val;
    std::vector<char> done;
    // callbacks
    std::map<ColumnRef, QM31> mask;
    std::vector<QM31> params;
    explicit GraphPointEval(const Graph& g_) : g(g_), val(g_.nodes.size()), done(g_.nodes.size(), 0) {}
    QM31 eval(int id) {
        if (done[id]) return val[id];
        const Node& n = g.nodes[id];
        QM31 r = qm_zero();
        switch (n.op) {
            case NodeOp::Col: {
                auto it = mask.find(ColumnRef{n.interaction, n.col, n.offset});
                if (it == mask.end()) throw std::logic_error("GraphPointEval: missing mask value");
                r = it->second;
                break;
            }
            case NodeOp::ConstF: r = qm_from_m31(n.fconst); break;
            case NodeOp::ConstE: r = n.econst; break;
            case NodeOp::ParamE: r = params.at(n.param); break;
            case NodeOp::AddF: case NodeOp::AddE: case NodeOp::AddEF: r = eval(n.a) + eval(n.b); break;
            case NodeOp::SubF: case NodeOp::SubE: case NodeOp::SubEF: r = eval(n.a) - eval(n.b); break;
            case NodeOp::MulF: case NodeOp::MulE: case NodeOp::MulEF: r = eval(n.a) * eval(n.b); break;
            case NodeOp::NegF: case NodeOp::NegE: r = -eval(n.a); break;
            case NodeOp::F2E: r = eval(n.a); break;
            case NodeOp::InvE: r = qm_inv(eval(n.a)); break;
            case NodeOp::InvF: case NodeOp::ShrF: case NodeOp::AndF: case NodeOp::RowLt: case NodeOp::LeF: case NodeOp::DivCF:
            case NodeOp::ModCF: case NodeOp::U32DivRemF:
                throw std::logic_error("GraphPointEval: witness-only op in a constraint graph");
            case NodeOp::Combine4:
                // combine_ef of QM31 "base" values (point.rs: from_partial_evals)
                r = qm_from_partial_evals(eval(n.a), eval(n.b), eval(n.c), eval(n.d));
                break;
        }
        val[id] = r;
        done[id] = 1;
        return r;
    }
};

// ------------------------------------------------------------------ compilation to bytecode
struct ProgramOutput {
    enum Kind { ConstraintSum, StoreE, StoreF, Hist } kind;
    int node;      // value node
    int slot = 0;  // ConstraintSum: index of the random-coefficient power param; Store*: out col; Hist: out col
};

class ProgramBuilder {
   public:
    // `col_index(interaction, col)` maps a graph column to an index into the kernel's in_cols table.
    // Extra QM31 parameters (beyond the evaluator's) can be appended by the caller: constraint
    // programs use params [n_eval_params + k] for the k-th random-coefficient power.
    template <class ColIndexFn>
    static AirProgram compile(const Graph& g, const std::vector<ProgramOutput>& outputs, size_t n_params, ColIndexFn col_index,
                              bool emit_cuda = false) {
        AirProgram prog;
        std::string& src = prog.cuda_body;
        auto F = [](int id) { return "f" + std::to_string(id); };
        auto E = [](int id) { return "e" + std::to_string(id); };
        auto U = [](u32 v) { return std::to_string(v) + "u"; };
        size_t n = g.nodes.size();
        // liveness
        std::vector<char> live(n, 0);
        std::vector<int> stack;
        for (auto& o : outputs) stack.push_back(o.node);
        while (!stack.empty()) {
            int id = stack.back();
            stack.pop_back();
            if (live[id]) continue;
            live[id] = 1;
            const Node& nd = g.nodes[id];
            for (int ch : {nd.a, nd.b, nd.c, nd.d})
                if (ch >= 0) stack.push_back(ch);
        }
        // outputs are emitted right after their value is computed: order them by node id
        std::vector<std::vector<size_t>> outs_at(n);
        for (size_t i = 0; i < outputs.size(); i++) outs_at[outputs[i].node].push_back(i);
        // last use (by node order; outputs count as a use at their own position)
        std::vector<int> last_use(n, -1);
        for (size_t id = 0; id < n; id++) {
            if (!live[id]) continue;
            const Node& nd = g.nodes[id];
            for (int ch : {nd.a, nd.b, nd.c, nd.d})
                if (ch >= 0) last_use[ch] = (int)id;
            if (last_use[id] < (int)id) last_use[id] = (int)id;
        }
        // ---- CUDA emission only: fuse sums of E*f products (LookupElements::combine, logup
        // numerators) into lazily reduced u64 dot products.  A chain  (((E0*f0 + E1*f1) + E2*f2) - z)
        // costs 4 multiply-adds per term + one reduction per coordinate instead of a reduced
        // multiplication and a reduced addition per term and coordinate.  Exact field arithmetic, so
        // the value equals the bytecode's; interior nodes must have no other user.
        struct FusedSum {
            std::vector<std::tuple<int, int, int>> terms;  // (sign, E node, F node)
            std::vector<std::pair<int, int>> addends;      // (sign, E node)
            std::vector<int> interior;
        };
        std::map<int, FusedSum> fused;
        std::vector<char> absorbed(n, 0);
        if (emit_cuda) {
            std::vector<int> uses(n, 0);
            for (size_t id = 0; id < n; id++) {
                if (!live[id]) continue;
                const Node& nd = g.nodes[id];
                for (int ch : {nd.a, nd.b, nd.c, nd.d})
                    if (ch >= 0) uses[ch]++;
            }
            for (auto& o : outputs) uses[o.node] += 2;  // outputs are never interior
            // Shared sub-sums (CSE across relation entries with a common prefix) are expanded into every
            // user: re-accumulating a term is 4 multiply-adds, cheaper than one reduced QM31 addition.
            // A shared node is still emitted on its own; nvcc drops it when every user was fused.
            std::function<void(int, int, FusedSum&, bool)> collect = [&](int id, int sign, FusedSum& f, bool root) {
                const Node& nd = g.nodes[id];
                bool single = root || uses[id] == 1;
                if (nd.op == NodeOp::MulEF && f.terms.size() < 48) {
                    f.terms.push_back(std::make_tuple(sign, nd.a, nd.b));
                    if (!root && single) f.interior.push_back(id);
                } else if ((nd.op == NodeOp::AddE || nd.op == NodeOp::SubE) && f.terms.size() < 48) {
                    if (!root && single) f.interior.push_back(id);
                    collect(nd.a, sign, f, false);
                    collect(nd.b, nd.op == NodeOp::AddE ? sign : -sign, f, false);
                } else {
                    f.addends.push_back({sign, id});
                }
            };
            for (int id = (int)n - 1; id >= 0; id--) {
                if (!live[id] || absorbed[id]) continue;
                const Node& nd = g.nodes[id];
                if (nd.op != NodeOp::AddE && nd.op != NodeOp::SubE) continue;
                FusedSum f;
                collect(id, 1, f, true);
                if (f.terms.size() < 2) continue;
                for (int i : f.interior) absorbed[i] = 1;
                fused[id] = f;
            }
        }
        // params table first (4 words each), then constants
        prog.param_slots.resize(n_params);
        for (size_t p = 0; p < n_params; p++) {
            prog.param_slots[p] = (u32)prog.consts.size();
            prog.consts.insert(prog.consts.end(), {0u, 0u, 0u, 0u});
        }
        std::map<u32, u32> fconst_slot;
        std::map<std::tuple<u32, u32, u32, u32>, u32> econst_slot;
        // register allocation: free lists for 1-word and 4-word (aligned) slots
        std::vector<u32> free1, free4;
        u32 next_reg = 0;
        auto alloc = [&](bool ext) -> u32 {
            if (ext) {
                if (!free4.empty()) {
                    u32 r = free4.back();
                    free4.pop_back();
                    return r;
                }
                u32 r = next_reg;
                next_reg += 4;
                return r;
            }
            if (free1.empty()) {
                if (!free4.empty()) {
                    u32 r = free4.back();
                    free4.pop_back();
                    for (u32 k = 0; k < 4; k++) free1.push_back(r + 3 - k);
                } else {
                    u32 r = next_reg;
                    next_reg += 4;
                    for (u32 k = 0; k < 4; k++) free1.push_back(r + 3 - k);
                }
            }
            u32 r = free1.back();
            free1.pop_back();
            return r;
        };
        std::vector<u32> reg(n, 0);
        std::vector<std::vector<int>> dying(n);  // nodes whose last use is at index i
        for (size_t id = 0; id < n; id++)
            if (live[id]) dying[last_use[id]].push_back((int)id);
        auto emit = [&](AirOp op, u32 dst, u32 a, u32 b) { prog.code.push_back(air_encode(op, dst, a, b)); };
        for (size_t id = 0; id < n; id++) {
            if (!live[id]) continue;
            const Node& nd = g.nodes[id];
            u32 r = alloc(nd.ext);
            reg[id] = r;
            const int nid = (int)id;
            auto fdef = [&](const std::string& rhs) { if (emit_cuda) src += "    const u32 " + F(nid) + " = " + rhs + ";\n"; };
            auto edef = [&](const std::string& rhs) {
                if (!emit_cuda || absorbed[nid]) return;
                auto fit = fused.find(nid);
                if (fit == fused.end()) {
                    src += "    const QM31 " + E(nid) + " = " + rhs + ";\n";
                    return;
                }
                const FusedSum& f = fit->second;
                std::string d = "d" + std::to_string(nid);
                src += "    u64 " + d + "_0 = 0, " + d + "_1 = 0, " + d + "_2 = 0, " + d + "_3 = 0;\n";
                size_t k = 0;
                for (auto& t : f.terms) {
                    std::string fv = std::get<0>(t) > 0 ? F(std::get<2>(t)) : "(P - " + F(std::get<2>(t)) + ")";
                    src += "    dot_term(" + d + "_0, " + d + "_1, " + d + "_2, " + d + "_3, " + E(std::get<1>(t)) + ", " + fv + ");\n";
                    if (++k % 4 == 0 && k != f.terms.size()) src += "    dot_fold(" + d + "_0, " + d + "_1, " + d + "_2, " + d + "_3);\n";
                }
                std::string expr = "dot_done(" + d + "_0, " + d + "_1, " + d + "_2, " + d + "_3)";
                for (auto& a : f.addends) expr = std::string(a.first > 0 ? "qm_add(" : "qm_sub(") + expr + ", " + E(a.second) + ")";
                src += "    const QM31 " + E(nid) + " = " + expr + ";\n";
            };
            switch (nd.op) {
                case NodeOp::Col: {
                    u32 ci = (u32)col_index(nd.interaction, nd.col);
                    emit(OP_LOAD, r, ci, (u32)(nd.offset & 0xfffff));
                    fdef(nd.offset == 0 ? "ldcol(" + U(ci) + ")" : "ldcol_off(" + U(ci) + ", " + std::to_string(nd.offset) + ")");
                    break;
                }
                case NodeOp::ConstF: {
                    auto it = fconst_slot.find(nd.fconst);
                    u32 s;
                    if (it == fconst_slot.end()) {
                        s = (u32)prog.consts.size();
                        prog.consts.push_back(nd.fconst);
                        fconst_slot[nd.fconst] = s;
                    } else s = it->second;
                    emit(OP_CONSTF, r, s, 0);
                    fdef("cw(" + U(s) + ")");
                    break;
                }
                case NodeOp::ConstE: {
                    auto key = std::make_tuple(nd.econst.a, nd.econst.b, nd.econst.c, nd.econst.d);
                    auto it = econst_slot.find(key);
                    u32 s;
                    if (it == econst_slot.end()) {
                        s = (u32)prog.consts.size();
                        prog.consts.insert(prog.consts.end(), {nd.econst.a, nd.econst.b, nd.econst.c, nd.econst.d});
                        econst_slot[key] = s;
                    } else s = it->second;
                    emit(OP_CONSTE, r, s, 0);
                    edef("cq(" + U(s) + ")");
                    break;
                }
                case NodeOp::ParamE:
                    emit(OP_CONSTE, r, prog.param_slots.at(nd.param), 0);
                    edef("cq(" + U(prog.param_slots.at(nd.param)) + ")");
                    break;
                case NodeOp::AddF: emit(OP_ADD, r, reg[nd.a], reg[nd.b]); fdef("m31_add(" + F(nd.a) + ", " + F(nd.b) + ")"); break;
                case NodeOp::SubF: emit(OP_SUB, r, reg[nd.a], reg[nd.b]); fdef("m31_sub(" + F(nd.a) + ", " + F(nd.b) + ")"); break;
                case NodeOp::MulF: emit(OP_MUL, r, reg[nd.a], reg[nd.b]); prog.n_mul_m31 += 1; fdef("m31_mul(" + F(nd.a) + ", " + F(nd.b) + ")"); break;
                case NodeOp::NegF: emit(OP_NEG, r, reg[nd.a], 0); fdef("m31_neg(" + F(nd.a) + ")"); break;
                case NodeOp::AddE: emit(OP_EADD, r, reg[nd.a], reg[nd.b]); edef("qm_add(" + E(nd.a) + ", " + E(nd.b) + ")"); break;
                case NodeOp::SubE: emit(OP_ESUB, r, reg[nd.a], reg[nd.b]); edef("qm_sub(" + E(nd.a) + ", " + E(nd.b) + ")"); break;
                case NodeOp::MulE: emit(OP_EMUL, r, reg[nd.a], reg[nd.b]); prog.n_mul_m31 += 16; edef("g_qm_mul(" + E(nd.a) + ", " + E(nd.b) + ")"); break;
                case NodeOp::NegE: emit(OP_ENEG, r, reg[nd.a], 0); edef("qm_neg(" + E(nd.a) + ")"); break;
                case NodeOp::MulEF: emit(OP_EMULF, r, reg[nd.a], reg[nd.b]); prog.n_mul_m31 += 4; edef("qm_mul_m31(" + E(nd.a) + ", " + F(nd.b) + ")"); break;
                case NodeOp::AddEF: emit(OP_EADDF, r, reg[nd.a], reg[nd.b]); edef("qm_add_m31(" + E(nd.a) + ", " + F(nd.b) + ")"); break;
                case NodeOp::SubEF: emit(OP_ESUBF, r, reg[nd.a], reg[nd.b]); edef("qm_sub_m31(" + E(nd.a) + ", " + F(nd.b) + ")"); break;
                case NodeOp::F2E: emit(OP_F2E, r, reg[nd.a], 0); edef("qm_from_m31(" + F(nd.a) + ")"); break;
                case NodeOp::InvE: emit(OP_EINV, r, reg[nd.a], 0); prog.n_mul_m31 += 100; edef("qm_inv(" + E(nd.a) + ")"); break;
                case NodeOp::InvF: emit(OP_INV, r, reg[nd.a], 0); prog.n_mul_m31 += 37; fdef("m31_inv(" + F(nd.a) + ")"); break;
                case NodeOp::ShrF: emit(OP_SHR, r, reg[nd.a], nd.fconst); fdef("(" + F(nd.a) + " >> " + U(nd.fconst) + ")"); break;
                case NodeOp::AndF: emit(OP_AND, r, reg[nd.a], nd.fconst); fdef("(" + F(nd.a) + " & " + U(nd.fconst) + ")"); break;
                case NodeOp::LeF: emit(OP_LE, r, reg[nd.a], reg[nd.b]); fdef("(" + F(nd.a) + " <= " + F(nd.b) + " ? 1u : 0u)"); break;
                case NodeOp::DivCF: emit(OP_DIVC, r, reg[nd.a], nd.fconst); fdef("(" + F(nd.a) + " / " + U(nd.fconst) + ")"); break;
                case NodeOp::ModCF: emit(OP_MODC, r, reg[nd.a], nd.fconst); fdef("(" + F(nd.a) + " % " + U(nd.fconst) + ")"); break;
                case NodeOp::U32DivRemF:
                    emit(OP_U32DIVREM, r, reg[nd.a], nd.fconst);
                    fdef("u32_divrem_part(" + E(nd.a) + ", " + U(nd.fconst) + ")");
                    break;
                case NodeOp::RowLt: {
                    u32 sl = (u32)prog.consts.size();
                    prog.consts.push_back(nd.fconst);
                    prog.rowlt_slots.push_back(sl);
                    emit(OP_ROWLT, r, sl, 0);
                    fdef("(row < cw(" + U(sl) + ") ? 1u : 0u)");
                    break;
                }
                case NodeOp::Combine4:
                    emit(OP_MOV, r, reg[nd.a], 0);
                    emit(OP_MOV, r + 1, reg[nd.b], 0);
                    emit(OP_MOV, r + 2, reg[nd.c], 0);
                    emit(OP_MOV, r + 3, reg[nd.d], 0);
                    edef("qm_make(" + F(nd.a) + ", " + F(nd.b) + ", " + F(nd.c) + ", " + F(nd.d) + ")");
                    break;
            }
            for (size_t oi : outs_at[id]) {
                const ProgramOutput& o = outputs[oi];
                switch (o.kind) {
                    case ProgramOutput::ConstraintSum:
                        emit(nd.ext ? OP_CONSTRAINT_E : OP_CONSTRAINT_F, 0, r, prog.param_slots.at(o.slot));
                        prog.n_mul_m31 += nd.ext ? 16 : 4;
                        if (emit_cuda)
                            // base-field constraints: coeff (QM31) x f accumulated as four lazily reduced u64 dot products
                            // (folded every 4 terms, reduced once at the end of the kernel) instead of 4 reduced multiplications
                            // and 4 reduced additions per constraint
                            src += nd.ext ? "    acc = qm_add(acc, g_qm_mul(cq(" + U(prog.param_slots.at(o.slot)) + "), " + E(nid) + "));\n"
                                          : "    CACC(cq(" + U(prog.param_slots.at(o.slot)) + "), " + F(nid) + ");\n";
                        break;
                    case ProgramOutput::StoreE:
                        emit(OP_STORE_E, 0, r, (u32)o.slot);
                        if (emit_cuda) src += "    st4(" + U((u32)o.slot) + ", " + E(nid) + ");\n";
                        break;
                    case ProgramOutput::StoreF:
                        emit(OP_STORE_F, 0, r, (u32)o.slot);
                        if (emit_cuda) src += "    st1(" + U((u32)o.slot) + ", " + F(nid) + ");\n";
                        break;
                    case ProgramOutput::Hist:
                        emit(OP_HIST, 0, r, (u32)o.slot);
                        if (emit_cuda) src += "    hist(" + U((u32)o.slot) + ", " + F(nid) + ");\n";
                        break;
                }
            }
            for (int d : dying[id]) {
                if (g.nodes[d].ext) free4.push_back(reg[d]);
                else free1.push_back(reg[d]);
            }
        }
        prog.n_regs = next_reg;
        return prog;
    }
};

}  // namespace cm31
