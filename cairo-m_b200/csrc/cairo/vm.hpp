// Host-side input producer for the proving hot path: a minimal Cairo-M VM for the felt opcode
// subset used by `fibonacci_loop`, and the adapter that turns its trace + memory log into the
// prover input (per-opcode ExecutionBundles + the global data-access log + boundary memory).
//
// These are the sequential steps immediately BEFORE the hot path; they are
// restated here only to feed the prover with the reference's input format:
//   VM step semantics      crates/runner/src/vm/instructions/{store,jnz,jump,call}.rs
//   memory trace order     crates/runner/src/memory/mod.rs:94-170 (instruction word(s) first, then operands)
//   ExecutionBundleIterator crates/prover/src/adapter/memory.rs:264-403
//   Memory::push            crates/prover/src/adapter/memory.rs:470-537 (prev clock/value, clock-update splitting)
//   update_multiplicities   crates/prover/src/adapter/memory.rs:411-457
//   import_internal         crates/prover/src/adapter/mod.rs:97-193
// Differences, by design: boundary-memory rows are emitted in ascending address order (the
// reference iterates two std HashMaps, i.e. in a per-run random order, components/memory.rs:105-109).
// The memory roots are Merkle roots under the reference's Poseidon2-M31 instance (csrc/cairo/poseidon2_constants.hpp).
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <vector>

#include "../air/cairo_components.hpp"

namespace cm31 {

struct Bundle {  // ExecutionBundle, flattened (12 words)
    u32 pc, fp, clock, inst_prev_clock;
    u32 inst[6];
    u32 span_start, span_len;
};
struct DataAccess {  // crates/prover/src/adapter/memory.rs:57-68
    u32 address, prev_clock, prev_value, value;
};
struct MemoryRow {  // one boundary-memory entry: (address, (value, clock, multiplicity))
    u32 address, clock;
    u32 value[4];
    u32 multiplicity, root;
};
struct ClockUpdateRow {
    u32 address, prev_clk;
    u32 value[4];
};
struct MerkleNode {  // adapter::merkle::NodeData::to_m31_array (crates/prover/src/adapter/merkle.rs:84-123) + the tree's root
    u32 index, depth, left_value, right_value, parent_value, left_multiplicity, right_multiplicity, parent_multiplicity, root;
};
struct PublicRanges {
    u32 program_start = 0, program_end = 0, input_start = 0, input_end = 0, output_start = 0, output_end = 0;
};
struct Registers {
    u32 pc, fp;
};

struct ProverInput {
    Registers initial_registers, final_registers;
    std::map<u32, std::vector<Bundle>> states_by_opcodes;
    std::vector<DataAccess> data_accesses;
    std::vector<MemoryRow> initial_memory, final_memory;  // ascending address
    std::vector<ClockUpdateRow> clock_update_data;
    PublicRanges public_ranges;
    u32 initial_root = 0, final_root = 0;
    std::vector<MerkleNode> merkle_nodes;  // initial tree then final tree (components/merkle.rs:82-104)
    size_t n_steps = 0;
};

struct Word4 {
    u32 v[4];
};

// ------------------------------------------------------------------ program: fibonacci_loop
// Hand-assembled CASM for test_data/functions/fibonacci_loop.cm (same opcode families as the
// compiler's while-loop lowering, crates/compiler/codegen/tests/snapshots/
// mdtest_codegen_snapshots@loops_in_cairo_m___while_loop.snap): 8 VM steps per loop iteration.
// Like the compiler's output, no instruction reads and writes the same cell (i = i + 1 goes
// through a temporary): two accesses of one cell in one step would need clock - prev_clock - 1 = -1
// in the RangeCheck20 relation.
inline std::vector<Word4> fibonacci_loop_program() {
    const u32 M3 = P - 3, M4 = P - 4, M8 = P - 8;
    return {
        {{OP_STORE_IMM, 0, 0, 0}},             //  0: a = [fp+0] = 0
        {{OP_STORE_IMM, 1, 1, 0}},             //  1: b = [fp+1] = 1
        {{OP_STORE_IMM, 0, 2, 0}},             //  2: i = [fp+2] = 0
        {{OP_STORE_SUB_FP_FP, 2, M4, 3}},      //  3: [fp+3] = i - n        (n = [fp-4])
        {{OP_JNZ_FP_IMM, 3, 2, 0}},            //  4: if [fp+3] != 0 jmp rel +2
        {{OP_JMP_REL_IMM, 7, 0, 0}},           //  5: jmp rel +7 (exit)
        {{OP_STORE_ADD_FP_FP, 0, 1, 4}},       //  6: temp = a + b
        {{OP_STORE_ADD_FP_IMM, 1, 0, 0}},      //  7: a = b + 0
        {{OP_STORE_ADD_FP_IMM, 4, 0, 1}},      //  8: b = temp + 0
        {{OP_STORE_ADD_FP_IMM, 2, 1, 5}},      //  9: [fp+5] = i + 1
        {{OP_STORE_ADD_FP_IMM, 5, 0, 2}},      // 10: i = [fp+5] + 0
        {{OP_JMP_REL_IMM, M8, 0, 0}},          // 11: jmp rel -8 (loop head)
        {{OP_STORE_ADD_FP_IMM, 0, 0, M3}},     // 12: [fp-3] = a + 0   (return value)
        {{OP_RET, 0, 0, 0}},                   // 13: ret
    };
}
inline size_t fibonacci_loop_steps(u32 n) { return 8 * (size_t)n + 8; }

// ------------------------------------------------------------------ program: array_sum
// Exercises the call / pointer opcode families: main fills arr[i] = i*i through a frame pointer
// (StoreFramePointer, StoreToDoubleDerefFpFp), patches arr[1] (StoreToDoubleDerefFpImm), calls
// sum(ptr, n) (CallAbsImm / Ret, StoreDoubleDerefFpFp), reads arr[0] back (StoreDoubleDerefFp) and
// asserts it is 0 (AssertEqFpImm).  Returns sum_{i<n} i^2 - 1 + n  (n >= 2).
// main frame: [fp-4] = n, [fp-3] = return slot; locals fp+0..7; callee frame at fp+8 (its fp = fp+10,
// args ptr = [fp'-4], n = [fp'-3], result [fp'-5]); the array starts at fp+20.
inline std::vector<Word4> array_sum_program() {
    const u32 M3 = P - 3, M4 = P - 4, M5 = P - 5, M8 = P - 8;
    return {
        {{OP_STORE_FRAME_POINTER, 20, 0, 0}},               //  0: ptr = fp + 20
        {{OP_STORE_IMM, 0, 1, 0}},                          //  1: i = 0
        {{OP_STORE_SUB_FP_FP, 1, M4, 2}},                   //  2: [fp+2] = i - n
        {{OP_JNZ_FP_IMM, 2, 2, 0}},                         //  3: if != 0 -> 5
        {{OP_JMP_REL_IMM, 7, 0, 0}},                        //  4: -> 11
        {{OP_STORE_ADD_FP_IMM, 1, 0, 5}},                   //  5: [fp+5] = i   (one step never touches a cell twice)
        {{OP_STORE_MUL_FP_FP, 1, 5, 3}},                    //  6: v = i * i
        {{OP_STORE_TO_DOUBLE_DEREF_FP_FP, 0, 1, 3}},        //  7: ptr[i] = v
        {{OP_STORE_ADD_FP_IMM, 1, 1, 4}},                   //  8: [fp+4] = i + 1
        {{OP_STORE_ADD_FP_IMM, 4, 0, 1}},                   //  9: i = [fp+4]
        {{OP_JMP_REL_IMM, M8, 0, 0}},                       // 10: -> 2
        {{OP_STORE_TO_DOUBLE_DEREF_FP_IMM, 0, 1, 1}},       // 11: ptr[1] = i (= n)
        {{OP_STORE_ADD_FP_IMM, 0, 0, 6}},                   // 12: arg ptr
        {{OP_STORE_ADD_FP_IMM, M4, 0, 7}},                  // 13: arg n
        {{OP_CALL_ABS_IMM, 8, 20, 0}},                      // 14: call sum (frame at fp+8)
        {{OP_STORE_DOUBLE_DEREF_FP, 0, 0, 4}},              // 15: [fp+4] = ptr[0]
        {{OP_ASSERT_EQ_FP_IMM, 4, 0, 0}},                   // 16: assert ptr[0] == 0
        {{OP_STORE_LE_FP_IMM, 1, 5, 2}},                    // 17: [fp+2] = (i <= 5), i = n here
        {{OP_STORE_ADD_FP_IMM, 5, 0, M3}},                  // 18: return value = sum (callee wrote [fp+5])
        {{OP_RET, 0, 0, 0}},                                // 19
        {{OP_STORE_IMM, 0, 0, 0}},                          // 20: sum: acc = 0
        {{OP_STORE_IMM, 0, 1, 0}},                          // 21: j = 0
        {{OP_STORE_SUB_FP_FP, 1, M3, 2}},                   // 22: [fp+2] = j - n
        {{OP_JNZ_FP_IMM, 2, 2, 0}},                         // 23: if != 0 -> 25
        {{OP_JMP_REL_IMM, 7, 0, 0}},                        // 24: -> 31
        {{OP_STORE_DOUBLE_DEREF_FP_FP, M4, 1, 3}},          // 25: x = ptr[j]
        {{OP_STORE_ADD_FP_FP, 0, 3, 4}},                    // 26: t = acc + x
        {{OP_STORE_ADD_FP_IMM, 4, 0, 0}},                   // 27: acc = t
        {{OP_STORE_ADD_FP_IMM, 1, 1, 5}},                   // 28: [fp+5] = j + 1
        {{OP_STORE_ADD_FP_IMM, 5, 0, 1}},                   // 29: j = [fp+5]
        {{OP_JMP_REL_IMM, M8, 0, 0}},                       // 30: -> 22
        {{OP_STORE_ADD_FP_IMM, 0, 0, M5}},                  // 31: result -> [fp-5]
        {{OP_RET, 0, 0, 0}},                                // 32
    };
}

// ------------------------------------------------------------------ program: u32_counter
// u32 arithmetic and bitwise ops: x = 0x0001fff0, y = 0x11; n times { t = x + y; x = (((t ^ y) & t) | y); y -= 1 }
// with wrap-around (y borrows through zero after 17 rounds, the first add carries across the limb
// boundary).  Returns x's low limb.
inline std::vector<Word4> u32_counter_program() {
    const u32 M3 = P - 3, M4 = P - 4;
    return {
        {{OP_U32_STORE_IMM, 0xfff0, 0x0001, 0}},        //  0: x
        {{OP_U32_STORE_IMM, 0x0011, 0x0000, 2}},        //  1: y
        {{OP_U32_STORE_IMM, 0, 0, 8}},                  //  2: zero
        {{OP_U32_STORE_IMM, 1, 0, 10}},                 //  3: one
        {{OP_STORE_IMM, 0, 6, 0}},                      //  4: i = 0
        {{OP_STORE_SUB_FP_FP, 6, M4, 7}},               //  5: [fp+7] = i - n
        {{OP_JNZ_FP_IMM, 7, 2, 0}},                     //  6: if != 0 -> 8
        {{OP_JMP_REL_IMM, 12, 0, 0}},                   //  7: -> 19
        {{OP_U32_STORE_ADD_FP_FP, 0, 2, 4}},            //  8: t = x + y
        {{OP_U32_STORE_XOR_FP_FP, 4, 2, 14}},           //  9: a = t ^ y
        {{OP_U32_STORE_AND_FP_FP, 14, 4, 16}},          // 10: b = a & t
        {{OP_U32_STORE_OR_FP_FP, 16, 2, 18}},           // 11: c = b | y
        {{OP_U32_STORE_ADD_FP_FP, 18, 8, 0}},           // 12: x = c + 0
        {{OP_U32_STORE_SUB_FP_FP, 2, 10, 12}},          // 13: t2 = y - 1
        {{OP_U32_STORE_ADD_FP_FP, 12, 8, 2}},           // 14: y = t2 + 0
        {{OP_U32_STORE_LT_FP_FP, 2, 0, 20}},            // 15: [fp+20] = (y < x)
        {{OP_STORE_ADD_FP_IMM, 6, 1, 13}},              // 16: [fp+13] = i + 1
        {{OP_STORE_ADD_FP_IMM, 13, 0, 6}},              // 17: i = [fp+13]
        {{OP_JMP_REL_IMM, P - 13, 0, 0}},               // 18: -> 5
        {{OP_STORE_ADD_FP_IMM, 0, 0, M3}},              // 19: return x.lo
        {{OP_RET, 0, 0, 0}},                            // 20
    };
}

// ------------------------------------------------------------------ program: u32_mix
// The u32 multiplication / division / comparison family and every two-word (*_fp_imm) u32 instruction: per round
//   t = x * y; u = t + K; (q, r) = u divrem y; v = q * M; w = v ^ A; a = w & B; b = a | 1; (c, d) = b divrem 7;
//   [fp+28] = (d < 3); [fp+0] = (c == q); x = b + r; y += 2
// 18 VM steps per round (18 n + 8 in total).  U32StoreEqFpFp writes [fp+0]: the reference's component only proves dst_off = 0
// (see U32StoreEqFpFpEval).  Two-word instructions occupy two consecutive cells; jump offsets count cells.
inline std::vector<Word4> u32_mix_program() {
    const u32 M3 = P - 3, M4 = P - 4;
    std::vector<Word4> p;
    auto one_word = [&](u32 op, u32 a, u32 b, u32 c) { p.push_back(Word4{{op, a, b, c}}); };
    auto two_words = [&](u32 op, u32 a, u32 b, u32 c, u32 d, u32 e = 0) {
        p.push_back(Word4{{op, a, b, c}});
        p.push_back(Word4{{d, e, 0, 0}});
    };
    one_word(OP_U32_STORE_IMM, 0x1234, 0x5678, 30);                      // x = 0x56781234
    one_word(OP_U32_STORE_IMM, 0x00f1, 0x0000, 2);                       // y = 0xf1
    one_word(OP_STORE_IMM, 0, 6, 0);                                     // i = 0
    const u32 loop = (u32)p.size();
    one_word(OP_STORE_SUB_FP_FP, 6, M4, 7);                              // [fp+7] = i - n
    one_word(OP_JNZ_FP_IMM, 7, 2, 0);                                    // if != 0 skip the exit jump
    const u32 exit_jmp = (u32)p.size();
    one_word(OP_JMP_REL_IMM, 0, 0, 0);                                   // -> exit (patched below)
    one_word(OP_U32_STORE_MUL_FP_FP, 30, 2, 8);                          // t = x * y
    two_words(OP_U32_STORE_ADD_FP_IMM, 8, 0x79b9, 0x9e37, 10);           // u = t + 0x9e3779b9
    two_words(OP_U32_STORE_DIV_REM_FP_FP, 10, 2, 12, 14);                // q = u / y, r = u % y
    two_words(OP_U32_STORE_MUL_FP_IMM, 12, 0x0065, 0x0001, 16);          // v = q * 0x10065
    two_words(OP_U32_STORE_XOR_FP_IMM, 16, 0xa5a5, 0x5a5a, 18);          // w = v ^ 0x5a5aa5a5
    two_words(OP_U32_STORE_AND_FP_IMM, 18, 0xffff, 0x0fff, 20);          // a = w & 0x0fffffff
    two_words(OP_U32_STORE_OR_FP_IMM, 20, 0x0001, 0x0000, 22);           // b = a | 1
    two_words(OP_U32_STORE_DIV_REM_FP_IMM, 22, 7, 0, 24, 26);            // c = b / 7, d = b % 7
    two_words(OP_U32_STORE_LT_FP_IMM, 26, 3, 0, 28);                     // [fp+28] = (d < 3)
    one_word(OP_U32_STORE_EQ_FP_FP, 24, 12, 0);                          // [fp+0] = (c == q)
    one_word(OP_U32_STORE_ADD_FP_FP, 22, 14, 30);                        // x = b + r
    two_words(OP_U32_STORE_ADD_FP_IMM, 2, 2, 0, 32);                     // y2 = y + 2
    two_words(OP_U32_STORE_ADD_FP_IMM, 32, 0, 0, 2);                     // y = y2
    one_word(OP_STORE_ADD_FP_IMM, 6, 1, 5);                              // [fp+5] = i + 1
    one_word(OP_STORE_ADD_FP_IMM, 5, 0, 6);                              // i = [fp+5]
    const u32 back = (u32)p.size();
    one_word(OP_JMP_REL_IMM, m31_sub(loop, back), 0, 0);                 // -> loop
    p[exit_jmp].v[1] = (u32)p.size() - exit_jmp;
    one_word(OP_STORE_ADD_FP_IMM, 30, 0, M3);                            // return x.lo
    one_word(OP_RET, 0, 0, 0);
    return p;
}

// SHA-256 (FIPS 180-4) the way examples/sha256-cairo-m/src/sha256.cm writes it -- BASELINE config 3, the u32 / bitwise /
// range-check heavy workload: rotr(x, n) = (x * 2^(32-n)) | (x / 2^n) through U32StoreMulFpImm + U32StoreDivRemFpImm +
// U32StoreOrFpFp (sha256.cm:17-28), Sigma / sigma / Ch / Maj through the byte-wise bitwise table (and, or, xor), the additions
// through the u32 limb adders.  n compressions of the padded block of "abc" chained through H (n = 1 gives sha256("abc"),
// the vector of the reference's own prover test, crates/prover/tests/prover.rs:247); one compression is straight-line code
// (48 schedule steps + 64 rounds, the working variables a..h renamed at assembly time instead of copied), ~3 490 VM steps.
// Returns the sum of the sixteen 16-bit limbs of H.  No instruction reads and writes the same cell.
inline std::vector<Word4> sha256_program() {
    static const u32 K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
        0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
        0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
        0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
        0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
        0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    static const u32 IV[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    const u32 M3 = P - 3, M4 = P - 4;
    std::vector<Word4> p;
    auto one_word = [&](u32 op, u32 a, u32 b, u32 c) { p.push_back(Word4{{op, a, b, c}}); };
    auto two_words = [&](u32 op, u32 a, u32 b, u32 c, u32 d, u32 e = 0) {
        p.push_back(Word4{{op, a, b, c}});
        p.push_back(Word4{{d, e, 0, 0}});
    };
    const u32 I = 0, I_NEXT = 1, CMP = 2, T0 = 200;  // frame: felts 0..2, H 10.., working variables 30.., W 50.., temporaries 200..
    auto H = [](u32 k) { return 10 + 2 * k; };
    auto S = [](u32 k) { return 30 + 2 * (k & 7); };
    auto W = [](u32 t) { return 50 + 2 * t; };
    u32 tnext = T0;
    auto T = [&]() {
        u32 t = tnext;
        tnext += 2;
        return t;
    };
    auto lo = [](u32 v) { return v & 0xffffu; };
    auto hi = [](u32 v) { return v >> 16; };
    auto imm = [&](u32 v, u32 dst) { one_word(OP_U32_STORE_IMM, lo(v), hi(v), dst); };
    auto bin_to = [&](u32 op, u32 a, u32 b, u32 dst) { one_word(op, a, b, dst); };
    auto bin = [&](u32 op, u32 a, u32 b) {
        u32 d = T();
        one_word(op, a, b, d);
        return d;
    };
    auto add = [&](u32 a, u32 b) { return bin(OP_U32_STORE_ADD_FP_FP, a, b); };
    auto band = [&](u32 a, u32 b) { return bin(OP_U32_STORE_AND_FP_FP, a, b); };
    auto bor = [&](u32 a, u32 b) { return bin(OP_U32_STORE_OR_FP_FP, a, b); };
    auto bxor = [&](u32 a, u32 b) { return bin(OP_U32_STORE_XOR_FP_FP, a, b); };
    auto addi = [&](u32 a, u32 v) {
        u32 d = T();
        two_words(OP_U32_STORE_ADD_FP_IMM, a, lo(v), hi(v), d);
        return d;
    };
    auto xori = [&](u32 a, u32 v) {
        u32 d = T();
        two_words(OP_U32_STORE_XOR_FP_IMM, a, lo(v), hi(v), d);
        return d;
    };
    auto copy = [&](u32 a, u32 dst) { two_words(OP_U32_STORE_ADD_FP_IMM, a, 0, 0, dst); };
    auto shr = [&](u32 a, u32 n) {  // x / 2^n (sha256.cm:40,45)
        u32 q = T(), r = T();
        two_words(OP_U32_STORE_DIV_REM_FP_IMM, a, lo(1u << n), hi(1u << n), q, r);
        return q;
    };
    auto rotr = [&](u32 a, u32 n) {  // sha256.cm:17-28
        u32 q = T(), r = T(), sh = T();
        two_words(OP_U32_STORE_DIV_REM_FP_IMM, a, lo(1u << n), hi(1u << n), q, r);
        two_words(OP_U32_STORE_MUL_FP_IMM, a, lo(1u << (32 - n)), hi(1u << (32 - n)), sh);
        return bor(sh, q);
    };
    for (u32 k = 0; k < 8; k++) imm(IV[k], H(k));
    one_word(OP_STORE_IMM, 0, I, 0);
    const u32 loop = (u32)p.size();
    one_word(OP_STORE_SUB_FP_FP, I, M4, CMP);  // i - n
    one_word(OP_JNZ_FP_IMM, CMP, 2, 0);
    const u32 exit_jmp = (u32)p.size();
    one_word(OP_JMP_REL_IMM, 0, 0, 0);  // -> exit (patched below)
    for (u32 k = 0; k < 8; k++) copy(H(k), S(k));
    imm(0x61626380u, W(0));  // "abc" + the 1 bit
    for (u32 t = 1; t < 15; t++) imm(0, W(t));
    imm(0x18, W(15));  // message length in bits
    for (u32 t = 16; t < 64; t++) {  // message schedule (sha256.cm:73-77)
        tnext = T0;
        u32 s0 = bxor(bxor(rotr(W(t - 15), 7), rotr(W(t - 15), 18)), shr(W(t - 15), 3));
        u32 s1 = bxor(bxor(rotr(W(t - 2), 17), rotr(W(t - 2), 19)), shr(W(t - 2), 10));
        bin_to(OP_U32_STORE_ADD_FP_FP, add(add(W(t - 16), s0), W(t - 7)), s1, W(t));
    }
    for (u32 t = 0; t < 64; t++) {  // compression rounds (sha256.cm:86-97); role r of round t lives in slot (r - t) mod 8
        tnext = T0;
        const u32 a = S(0 - t), b = S(1 - t), c = S(2 - t), d = S(3 - t), e = S(4 - t), f = S(5 - t), g = S(6 - t), h = S(7 - t);
        u32 s1 = bxor(bxor(rotr(e, 6), rotr(e, 11)), rotr(e, 25));
        u32 ch = bxor(band(e, f), band(xori(e, 0xffffffffu), g));
        u32 t1 = add(addi(add(add(h, s1), ch), K[t]), W(t));
        u32 s0 = bxor(bxor(rotr(a, 2), rotr(a, 13)), rotr(a, 22));
        u32 maj = bxor(bxor(band(a, b), band(a, c)), band(b, c));
        u32 t2 = add(s0, maj);
        copy(add(d, t1), d);                                  // e' = d + temp1, in d's slot (= role e of round t + 1)
        bin_to(OP_U32_STORE_ADD_FP_FP, t1, t2, h);            // a' = temp1 + temp2, in h's slot (= role a of round t + 1)
    }
    tnext = T0;
    for (u32 k = 0; k < 8; k++) copy(add(H(k), S(k)), H(k));  // after 64 rounds role k is back in slot k
    one_word(OP_STORE_ADD_FP_IMM, I, 1, I_NEXT);
    one_word(OP_STORE_ADD_FP_IMM, I_NEXT, 0, I);
    const u32 back = (u32)p.size();
    one_word(OP_JMP_REL_IMM, m31_sub(loop, back), 0, 0);
    p[exit_jmp].v[1] = (u32)p.size() - exit_jmp;
    u32 acc = H(0);  // felt sum of the sixteen limbs
    for (u32 k = 1; k < 16; k++) {
        u32 dst = T0 + (k & 1);
        one_word(OP_STORE_ADD_FP_FP, acc, H(0) + k, dst);
        acc = dst;
    }
    one_word(OP_STORE_ADD_FP_IMM, acc, 0, M3);
    one_word(OP_RET, 0, 0, 0);
    return p;
}

// ------------------------------------------------------------------ program: all_opcodes (BASELINE config 4)
// The synthetic "all components" workload: every loop iteration executes every opcode family once or more, so after n
// iterations 25 of the 26 opcode components hold n .. 4n live rows each (u32_store_eq_fp_imm can only prove padding: see
// U32StoreEqFpImmEval).  One iteration = a felt block (store / add / sub / mul / div, le, assert), a pointer block (frame
// pointer, the four double-deref forms), a call / ret pair, an absolute jump, the u32_mix round (u32 mul / divrem / eq / lt and
// every two-word *_fp_imm instruction) and the u32_counter round (u32 add / sub / and / or / xor / lt on registers, one u32 constant): 46 VM
// steps.  Returns the low limb of u32_mix's x (same recurrence as u32_mix_program).
inline std::vector<Word4> all_opcodes_program() {
    const u32 M3 = P - 3, M4 = P - 4;
    std::vector<Word4> p;
    auto one_word = [&](u32 op, u32 a, u32 b, u32 c) { p.push_back(Word4{{op, a, b, c}}); };
    auto two_words = [&](u32 op, u32 a, u32 b, u32 c, u32 d, u32 e = 0) {
        p.push_back(Word4{{op, a, b, c}});
        p.push_back(Word4{{d, e, 0, 0}});
    };
    // felt slots
    const u32 EQ = 0, I = 1, I_NEXT = 2, CMP = 3, A = 4, B = 5, T = 6, U = 7, V = 8, W = 9, D = 10, LE = 11, PTR = 12, IDX = 13, LD1 = 14, LD2 = 15;
    const u32 ARRAY = 80, CALLEE = 90;  // ptr = fp + 80; the callee's frame starts at fp + 90 (its fp = fp + 92)
    const u32 MX = 100, CN = 140;       // u32 registers of the u32_mix / u32_counter rounds
    one_word(OP_U32_STORE_IMM, 0x1234, 0x5678, MX + 30);  // mix: x = 0x56781234
    one_word(OP_U32_STORE_IMM, 0x00f1, 0x0000, MX + 2);   // mix: y = 0xf1
    one_word(OP_U32_STORE_IMM, 0xfff0, 0x0001, CN + 0);   // counter: x
    one_word(OP_U32_STORE_IMM, 0x0011, 0x0000, CN + 2);   // counter: y
    one_word(OP_U32_STORE_IMM, 0, 0, CN + 8);             // counter: zero
    one_word(OP_U32_STORE_IMM, 1, 0, CN + 10);            // counter: one
    one_word(OP_STORE_IMM, 0, I, 0);
    const u32 loop = (u32)p.size();
    one_word(OP_STORE_SUB_FP_FP, I, M4, CMP);  // i - n
    one_word(OP_JNZ_FP_IMM, CMP, 2, 0);
    const u32 exit_jmp = (u32)p.size();
    one_word(OP_JMP_REL_IMM, 0, 0, 0);  // -> exit (patched)
    // ---- felt block
    one_word(OP_STORE_IMM, 7, A, 0);               // a = 7
    one_word(OP_STORE_ADD_FP_IMM, A, 3, B);        // b = a + 3
    one_word(OP_STORE_MUL_FP_IMM, B, 5, T);        // t = b * 5
    one_word(OP_STORE_ADD_FP_FP, A, B, U);         // u = a + b
    one_word(OP_STORE_SUB_FP_FP, T, U, V);         // v = t - u
    one_word(OP_STORE_MUL_FP_FP, V, U, W);         // w = v * u
    one_word(OP_STORE_DIV_FP_FP, W, B, D);         // d = w / b
    one_word(OP_STORE_LE_FP_IMM, A, 9, LE);        // le = (a <= 9)
    one_word(OP_ASSERT_EQ_FP_IMM, A, 7, 0);        // assert a == 7
    // ---- pointer block
    one_word(OP_STORE_FRAME_POINTER, ARRAY, PTR, 0);           // ptr = fp + 80
    one_word(OP_STORE_IMM, 2, IDX, 0);                         // idx = 2
    one_word(OP_STORE_TO_DOUBLE_DEREF_FP_IMM, PTR, 1, A);      // ptr[1] = a
    one_word(OP_STORE_TO_DOUBLE_DEREF_FP_FP, PTR, IDX, B);     // ptr[idx] = b
    one_word(OP_STORE_DOUBLE_DEREF_FP, PTR, 1, LD1);           // ld1 = ptr[1]
    one_word(OP_STORE_DOUBLE_DEREF_FP_FP, PTR, IDX, LD2);      // ld2 = ptr[idx]
    // ---- call / ret, absolute jump
    const u32 call_at = (u32)p.size();
    one_word(OP_CALL_ABS_IMM, CALLEE, 0, 0);                   // target patched below
    one_word(OP_JMP_ABS_IMM, (u32)p.size() + 1, 0, 0);         // jmp abs -> the next instruction
    // ---- the u32_mix round (registers at fp + 100 ..; the equality lands in [fp+0], see U32StoreEqFpFpEval)
    one_word(OP_U32_STORE_MUL_FP_FP, MX + 30, MX + 2, MX + 8);                  // t = x * y
    two_words(OP_U32_STORE_ADD_FP_IMM, MX + 8, 0x79b9, 0x9e37, MX + 10);        // u = t + 0x9e3779b9
    two_words(OP_U32_STORE_DIV_REM_FP_FP, MX + 10, MX + 2, MX + 12, MX + 14);   // q = u / y, r = u % y
    two_words(OP_U32_STORE_MUL_FP_IMM, MX + 12, 0x0065, 0x0001, MX + 16);       // v = q * 0x10065
    two_words(OP_U32_STORE_XOR_FP_IMM, MX + 16, 0xa5a5, 0x5a5a, MX + 18);       // w = v ^ 0x5a5aa5a5
    two_words(OP_U32_STORE_AND_FP_IMM, MX + 18, 0xffff, 0x0fff, MX + 20);       // a = w & 0x0fffffff
    two_words(OP_U32_STORE_OR_FP_IMM, MX + 20, 0x0001, 0x0000, MX + 22);        // b = a | 1
    two_words(OP_U32_STORE_DIV_REM_FP_IMM, MX + 22, 7, 0, MX + 24, MX + 26);    // c = b / 7, d = b % 7
    two_words(OP_U32_STORE_LT_FP_IMM, MX + 26, 3, 0, MX + 28);                  // (d < 3)
    one_word(OP_U32_STORE_EQ_FP_FP, MX + 24, MX + 12, EQ);                      // [fp+0] = (c == q)
    one_word(OP_U32_STORE_ADD_FP_FP, MX + 22, MX + 14, MX + 30);                // x = b + r
    two_words(OP_U32_STORE_ADD_FP_IMM, MX + 2, 2, 0, MX + 32);                  // y2 = y + 2
    two_words(OP_U32_STORE_ADD_FP_IMM, MX + 32, 0, 0, MX + 2);                  // y = y2
    // ---- the u32_counter round (registers at fp + 140 ..)
    one_word(OP_U32_STORE_ADD_FP_FP, CN + 0, CN + 2, CN + 4);                   // t = x + y
    one_word(OP_U32_STORE_XOR_FP_FP, CN + 4, CN + 2, CN + 14);                  // a = t ^ y
    one_word(OP_U32_STORE_AND_FP_FP, CN + 14, CN + 4, CN + 16);                 // b = a & t
    one_word(OP_U32_STORE_OR_FP_FP, CN + 16, CN + 2, CN + 18);                  // c = b | y
    one_word(OP_U32_STORE_ADD_FP_FP, CN + 18, CN + 8, CN + 0);                  // x = c + 0
    one_word(OP_U32_STORE_SUB_FP_FP, CN + 2, CN + 10, CN + 12);                 // t2 = y - 1
    one_word(OP_U32_STORE_ADD_FP_FP, CN + 12, CN + 8, CN + 2);                  // y = t2 + 0
    one_word(OP_U32_STORE_LT_FP_FP, CN + 2, CN + 0, CN + 20);                   // (y < x)
    one_word(OP_U32_STORE_IMM, 0x00ff, 0x0000, CN + 22);                        // a scratch u32 constant (u32_store_imm live)
    // ---- loop control
    one_word(OP_STORE_ADD_FP_IMM, I, 1, I_NEXT);
    one_word(OP_STORE_ADD_FP_IMM, I_NEXT, 0, I);
    const u32 back = (u32)p.size();
    one_word(OP_JMP_REL_IMM, m31_sub(loop, back), 0, 0);
    p[exit_jmp].v[1] = (u32)p.size() - exit_jmp;
    one_word(OP_STORE_ADD_FP_IMM, MX + 30, 0, M3);  // return mix x.lo
    one_word(OP_RET, 0, 0, 0);
    p[call_at].v[2] = (u32)p.size();                // callee: one store, ret
    one_word(OP_STORE_IMM, 5, 0, 0);
    one_word(OP_RET, 0, 0, 0);
    return p;
}

enum CairoProgramId : u32 {
    PROGRAM_FIBONACCI_LOOP = 0, PROGRAM_ARRAY_SUM = 1, PROGRAM_U32_COUNTER = 2, PROGRAM_U32_MIX = 3, PROGRAM_SHA256 = 4, PROGRAM_ALL_OPCODES = 5
};
inline std::vector<Word4> program_by_id(u32 id) {
    switch (id) {
        case PROGRAM_FIBONACCI_LOOP: return fibonacci_loop_program();
        case PROGRAM_ARRAY_SUM: return array_sum_program();
        case PROGRAM_U32_COUNTER: return u32_counter_program();
        case PROGRAM_U32_MIX: return u32_mix_program();
        case PROGRAM_SHA256: return sha256_program();
        case PROGRAM_ALL_OPCODES: return all_opcodes_program();
        default: throw std::runtime_error("unknown program id");
    }
}

// ------------------------------------------------------------------ VM
inline int opcode_size_in_m31s(u32 op);
struct VmTrace {
    std::vector<Registers> trace;                        // one entry per step + the final state
    std::vector<std::pair<u32, Word4>> memory_trace;     // (addr, value) in access order
    std::vector<Word4> initial_memory;                   // dense, address = index
    PublicRanges public_ranges;
    u32 return_value = 0;
};

// Runs `program` with one felt argument: frame = [arg, ret slot, old fp, return pc], fp after it.
// `segments` + `segment_steps`: continuation segments as the reference runner cuts them (crates/runner/src/vm/mod.rs:158-285,
// RunnerOptions::max_steps): a segment ends once its trace holds `segment_steps` states, gets the current state appended as
// its final entry, and the next segment starts from that state with the whole memory image as its initial memory (clocks
// restart).  The segments before the last one are appended to *segments; the last one is the return value.
inline VmTrace run_program(const std::vector<Word4>& program, u32 arg, size_t max_steps = (size_t)1 << 30,
                           std::vector<VmTrace>* segments = nullptr, size_t segment_steps = 0) {
    VmTrace out;
    u32 prog_len = (u32)program.size();
    std::vector<Word4> mem(program);
    auto cell = [&](u32 addr) -> Word4& {
        if (addr >= (1u << 30)) throw std::runtime_error("vm: address out of bounds");
        if (addr >= mem.size()) mem.resize(addr + 1, Word4{{0, 0, 0, 0}});
        return mem[addr];
    };
    u32 fp0 = prog_len + 4;
    cell(prog_len + 0) = Word4{{arg % P, 0, 0, 0}};
    cell(prog_len + 1) = Word4{{0, 0, 0, 0}};
    cell(prog_len + 2) = Word4{{fp0, 0, 0, 0}};       // old fp
    cell(prog_len + 3) = Word4{{prog_len, 0, 0, 0}};  // return pc = end of program
    out.initial_memory = mem;
    out.public_ranges.program_start = 0;
    out.public_ranges.program_end = prog_len;
    out.public_ranges.input_start = prog_len;
    out.public_ranges.input_end = prog_len + 1;
    out.public_ranges.output_start = prog_len + 1;
    out.public_ranges.output_end = prog_len + 2;
    u32 pc = 0, fp = fp0;
    auto rd = [&](u32 addr) -> u32 {
        Word4 w = cell(addr);
        out.memory_trace.push_back({addr, w});
        return w.v[0];
    };
    auto wr = [&](u32 addr, u32 val) {
        Word4 w{{val, 0, 0, 0}};
        cell(addr) = w;
        out.memory_trace.push_back({addr, w});
    };
    size_t steps = 0;
    while (pc != prog_len) {
        if (steps++ >= max_steps) throw std::runtime_error("vm: step limit reached");
        if (segments && segment_steps && out.trace.size() >= segment_steps) {  // ExecutionStatus::Ongoing -> finalize_segment(false)
            out.trace.push_back(Registers{pc, fp});
            VmTrace next;
            next.public_ranges = out.public_ranges;
            next.initial_memory = mem;
            segments->push_back(std::move(out));
            out = std::move(next);
        }
        out.trace.push_back(Registers{pc, fp});
        if (pc >= prog_len) throw std::runtime_error("vm: pc outside the program");
        Word4 ins = mem[pc];
        out.memory_trace.push_back({pc, ins});
        u32 op = ins.v[0], a = ins.v[1], b = ins.v[2], c = ins.v[3];
        int size_m31 = opcode_size_in_m31s(op);
        u32 d = 0, e = 0;  // operands 4 and 5 live in the second QM31 word of a two-word instruction
        if (size_m31 > 4) {
            if (pc + 1 >= prog_len) throw std::runtime_error("vm: truncated two-word instruction");
            Word4 ins2 = mem[pc + 1];
            out.memory_trace.push_back({pc + 1, ins2});
            d = ins2.v[0];
            e = ins2.v[1];
        }
        const u32 pc_step = size_m31 > 4 ? 2 : 1;  // State::advance_by(size_in_qm31s)
        switch (op) {
            case OP_STORE_ADD_FP_FP: case OP_STORE_SUB_FP_FP: case OP_STORE_MUL_FP_FP: case OP_STORE_DIV_FP_FP: {
                u32 x = rd(m31_add(fp, a)), y = rd(m31_add(fp, b));
                u32 r = op == OP_STORE_ADD_FP_FP ? m31_add(x, y) : op == OP_STORE_SUB_FP_FP ? m31_sub(x, y)
                        : op == OP_STORE_MUL_FP_FP ? m31_mul(x, y) : m31_mul(x, m31_inv(y));
                wr(m31_add(fp, c), r);
                pc += 1;
                break;
            }
            case OP_STORE_ADD_FP_IMM: case OP_STORE_MUL_FP_IMM: {
                u32 x = rd(m31_add(fp, a));
                wr(m31_add(fp, c), op == OP_STORE_ADD_FP_IMM ? m31_add(x, b) : m31_mul(x, b));
                pc += 1;
                break;
            }
            case OP_STORE_IMM:
                wr(m31_add(fp, b), a);
                pc += 1;
                break;
            case OP_JNZ_FP_IMM: {
                u32 cond = rd(m31_add(fp, a));
                pc = cond != 0 ? m31_add(pc, b) : pc + 1;
                break;
            }
            case OP_CALL_ABS_IMM: {  // call.rs:48-61: [fp+off0] = fp, [fp+off0+1] = pc+1, fp += off0+2, pc = target
                wr(m31_add(fp, a), fp);
                wr(m31_add(m31_add(fp, a), 1), pc + 1);
                fp = m31_add(m31_add(fp, a), 2);
                pc = b;
                break;
            }
            case OP_ASSERT_EQ_FP_IMM: {  // assert.rs:13-26
                if (rd(m31_add(fp, a)) != b) throw std::runtime_error("vm: assertion failed");
                pc += 1;
                break;
            }
            case OP_U32_STORE_IMM:  // store.rs:431-449: limbs imm_lo, imm_hi at [fp+dst], [fp+dst+1]
                if (a > 0xffff || b > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                wr(m31_add(fp, c), a);
                wr(m31_add(m31_add(fp, c), 1), b);
                pc += 1;
                break;
            case OP_U32_STORE_AND_FP_FP: case OP_U32_STORE_OR_FP_FP: case OP_U32_STORE_XOR_FP_FP:
            case OP_U32_STORE_ADD_FP_FP: case OP_U32_STORE_SUB_FP_FP: {  // exec_u32_bin_op_fp_fp (store.rs:15-35)
                u32 x_lo = rd(m31_add(fp, a)), x_hi = rd(m31_add(m31_add(fp, a), 1));
                u32 y_lo = rd(m31_add(fp, b)), y_hi = rd(m31_add(m31_add(fp, b), 1));
                if ((x_lo | x_hi | y_lo | y_hi) > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                u32 x = (x_hi << 16) | x_lo, y = (y_hi << 16) | y_lo;
                u32 r = op == OP_U32_STORE_ADD_FP_FP ? x + y : op == OP_U32_STORE_SUB_FP_FP ? x - y : op == OP_U32_STORE_AND_FP_FP ? (x & y)
                        : op == OP_U32_STORE_OR_FP_FP ? (x | y) : (x ^ y);
                wr(m31_add(fp, c), r & 0xffff);
                wr(m31_add(m31_add(fp, c), 1), r >> 16);
                pc += 1;
                break;
            }
            case OP_U32_STORE_LT_FP_FP: {  // [fp+dst] = u32(src0) < u32(src1)
                u32 x_lo = rd(m31_add(fp, a)), x_hi = rd(m31_add(m31_add(fp, a), 1));
                u32 y_lo = rd(m31_add(fp, b)), y_hi = rd(m31_add(m31_add(fp, b), 1));
                if ((x_lo | x_hi | y_lo | y_hi) > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                wr(m31_add(fp, c), ((x_hi << 16) | x_lo) < ((y_hi << 16) | y_lo) ? 1u : 0u);
                pc += 1;
                break;
            }
            case OP_U32_STORE_MUL_FP_FP: {  // wrapping_mul (store.rs:343-344)
                u32 x_lo = rd(m31_add(fp, a)), x_hi = rd(m31_add(m31_add(fp, a), 1));
                u32 y_lo = rd(m31_add(fp, b)), y_hi = rd(m31_add(m31_add(fp, b), 1));
                if ((x_lo | x_hi | y_lo | y_hi) > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                u32 r = ((x_hi << 16) | x_lo) * ((y_hi << 16) | y_lo);
                wr(m31_add(fp, c), r & 0xffff);
                wr(m31_add(m31_add(fp, c), 1), r >> 16);
                pc += 1;
                break;
            }
            case OP_U32_STORE_EQ_FP_FP: {  // [fp+dst] = u32(src0) == u32(src1)
                u32 x_lo = rd(m31_add(fp, a)), x_hi = rd(m31_add(m31_add(fp, a), 1));
                u32 y_lo = rd(m31_add(fp, b)), y_hi = rd(m31_add(m31_add(fp, b), 1));
                if ((x_lo | x_hi | y_lo | y_hi) > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                wr(m31_add(fp, c), (x_lo == y_lo && x_hi == y_hi) ? 1u : 0u);
                pc += 1;
                break;
            }
            case OP_U32_STORE_DIV_REM_FP_FP: {  // u32_store_div_rem_fp_fp (store.rs:346-373): dst = n / d, dst_rem = n % d
                u32 x_lo = rd(m31_add(fp, a)), x_hi = rd(m31_add(m31_add(fp, a), 1));
                u32 y_lo = rd(m31_add(fp, b)), y_hi = rd(m31_add(m31_add(fp, b), 1));
                if ((x_lo | x_hi | y_lo | y_hi) > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                u32 x = (x_hi << 16) | x_lo, y = (y_hi << 16) | y_lo;
                if (y == 0) throw std::runtime_error("vm: division by zero");
                wr(m31_add(fp, c), (x / y) & 0xffff);
                wr(m31_add(m31_add(fp, c), 1), (x / y) >> 16);
                wr(m31_add(fp, d), (x % y) & 0xffff);
                wr(m31_add(m31_add(fp, d), 1), (x % y) >> 16);
                pc += pc_step;
                break;
            }
            case OP_U32_STORE_ADD_FP_IMM: case OP_U32_STORE_MUL_FP_IMM: case OP_U32_STORE_AND_FP_IMM: case OP_U32_STORE_OR_FP_IMM:
            case OP_U32_STORE_XOR_FP_IMM: {  // exec_u32_bin_op_fp_imm (store.rs:37-66): operands src_off, imm_lo, imm_hi, dst_off
                if (b > 0xffff || c > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                u32 x_lo = rd(m31_add(fp, a)), x_hi = rd(m31_add(m31_add(fp, a), 1));
                if ((x_lo | x_hi) > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                u32 x = (x_hi << 16) | x_lo, y = (c << 16) | b;
                u32 r = op == OP_U32_STORE_ADD_FP_IMM ? x + y : op == OP_U32_STORE_MUL_FP_IMM ? x * y : op == OP_U32_STORE_AND_FP_IMM ? (x & y)
                        : op == OP_U32_STORE_OR_FP_IMM ? (x | y) : (x ^ y);
                wr(m31_add(fp, d), r & 0xffff);
                wr(m31_add(m31_add(fp, d), 1), r >> 16);
                pc += pc_step;
                break;
            }
            case OP_U32_STORE_LT_FP_IMM: {  // [fp+dst] = u32(src) < u32(imm)
                if (b > 0xffff || c > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                u32 x_lo = rd(m31_add(fp, a)), x_hi = rd(m31_add(m31_add(fp, a), 1));
                if ((x_lo | x_hi) > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                wr(m31_add(fp, d), ((x_hi << 16) | x_lo) < ((c << 16) | b) ? 1u : 0u);
                pc += pc_step;
                break;
            }
            case OP_U32_STORE_DIV_REM_FP_IMM: {  // u32_store_div_rem_fp_imm (store.rs:375-419): operands src, imm_lo, imm_hi, dst, dst_rem
                if (b > 0xffff || c > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                u32 y = (c << 16) | b;
                if (y == 0) throw std::runtime_error("vm: division by zero");
                u32 x_lo = rd(m31_add(fp, a)), x_hi = rd(m31_add(m31_add(fp, a), 1));
                if ((x_lo | x_hi) > 0xffff) throw std::runtime_error("vm: u32 limb out of range");
                u32 x = (x_hi << 16) | x_lo;
                wr(m31_add(fp, d), (x / y) & 0xffff);
                wr(m31_add(m31_add(fp, d), 1), (x / y) >> 16);
                wr(m31_add(fp, e), (x % y) & 0xffff);
                wr(m31_add(m31_add(fp, e), 1), (x % y) >> 16);
                pc += pc_step;
                break;
            }
            case OP_STORE_LE_FP_IMM: {  // [fp+dst] = ([fp+src] <= imm)   (store.rs:179-191)
                u32 x = rd(m31_add(fp, a));
                wr(m31_add(fp, c), x <= b ? 1u : 0u);
                pc += 1;
                break;
            }
            case OP_STORE_FRAME_POINTER:  // [fp+dst_off] = fp + imm
                wr(m31_add(fp, b), m31_add(fp, a));
                pc += 1;
                break;
            case OP_STORE_DOUBLE_DEREF_FP: {  // [fp+dst] = [[fp+base]+imm]   (store.rs:201-213)
                u32 base = rd(m31_add(fp, a));
                u32 v = rd(m31_add(base, b));
                wr(m31_add(fp, c), v);
                pc += 1;
                break;
            }
            case OP_STORE_TO_DOUBLE_DEREF_FP_IMM: {  // [[fp+base]+imm] = [fp+src]   (store.rs:258-275)
                u32 base = rd(m31_add(fp, a));
                u32 v = rd(m31_add(fp, c));
                wr(m31_add(base, b), v);
                pc += 1;
                break;
            }
            case OP_STORE_DOUBLE_DEREF_FP_FP: {  // [fp+dst] = [[fp+base]+[fp+offset]]   (store.rs:219-245)
                u32 base = rd(m31_add(fp, a));
                u32 off = rd(m31_add(fp, b));
                u32 v = rd(m31_add(base, off));
                wr(m31_add(fp, c), v);
                pc += 1;
                break;
            }
            case OP_STORE_TO_DOUBLE_DEREF_FP_FP: {  // [[fp+base]+[fp+offset]] = [fp+src]   (store.rs:284-303)
                u32 base = rd(m31_add(fp, a));
                u32 off = rd(m31_add(fp, b));
                u32 v = rd(m31_add(fp, c));
                wr(m31_add(base, off), v);
                pc += 1;
                break;
            }
            case OP_JMP_ABS_IMM: pc = a; break;
            case OP_JMP_REL_IMM: pc = m31_add(pc, a); break;
            case OP_RET: {
                u32 new_pc = rd(m31_sub(fp, 1));
                u32 new_fp = rd(m31_sub(fp, 2));
                pc = new_pc;
                fp = new_fp;
                break;
            }
            default: throw std::runtime_error("vm: unsupported opcode");
        }
    }
    out.trace.push_back(Registers{pc, fp});
    out.return_value = mem[prog_len + 1].v[0];
    return out;
}

inline int opcode_memory_accesses(u32 op) {
    switch (op) {
        case OP_STORE_ADD_FP_FP: case OP_STORE_SUB_FP_FP: case OP_STORE_MUL_FP_FP: case OP_STORE_DIV_FP_FP: return 3;
        case OP_STORE_ADD_FP_IMM: case OP_STORE_MUL_FP_IMM: case OP_STORE_LE_FP_IMM: return 2;
        case OP_STORE_IMM: return 1;
        case OP_JNZ_FP_IMM: return 1;
        case OP_JMP_ABS_IMM: case OP_JMP_REL_IMM: return 0;
        case OP_RET: return 2;
        case OP_CALL_ABS_IMM: case OP_U32_STORE_IMM: return 2;
        case OP_U32_STORE_LT_FP_FP: case OP_U32_STORE_EQ_FP_FP: return 5;
        case OP_U32_STORE_MUL_FP_FP: return 6;
        case OP_U32_STORE_DIV_REM_FP_FP: return 8;
        case OP_U32_STORE_ADD_FP_IMM: case OP_U32_STORE_MUL_FP_IMM: case OP_U32_STORE_AND_FP_IMM: case OP_U32_STORE_OR_FP_IMM:
        case OP_U32_STORE_XOR_FP_IMM: return 4;
        case OP_U32_STORE_LT_FP_IMM: return 3;
        case OP_U32_STORE_DIV_REM_FP_IMM: return 6;
        case OP_U32_STORE_ADD_FP_FP: case OP_U32_STORE_SUB_FP_FP: case OP_U32_STORE_AND_FP_FP: case OP_U32_STORE_OR_FP_FP:
        case OP_U32_STORE_XOR_FP_FP: return 6;
        case OP_ASSERT_EQ_FP_IMM: case OP_STORE_FRAME_POINTER: return 1;
        case OP_STORE_DOUBLE_DEREF_FP: case OP_STORE_TO_DOUBLE_DEREF_FP_IMM: return 3;
        case OP_STORE_DOUBLE_DEREF_FP_FP: case OP_STORE_TO_DOUBLE_DEREF_FP_FP: return 4;
        default: return -1;
    }
}
inline int opcode_size_in_m31s(u32 op) {
    switch (op) {
        case OP_RET: return 1;
        case OP_JMP_ABS_IMM: case OP_JMP_REL_IMM: return 2;
        case OP_STORE_IMM: case OP_JNZ_FP_IMM: case OP_CALL_ABS_IMM: case OP_ASSERT_EQ_FP_IMM: case OP_STORE_FRAME_POINTER: return 3;
        case OP_U32_STORE_DIV_REM_FP_FP: case OP_U32_STORE_ADD_FP_IMM: case OP_U32_STORE_MUL_FP_IMM: case OP_U32_STORE_LT_FP_IMM:
        case OP_U32_STORE_AND_FP_IMM: case OP_U32_STORE_OR_FP_IMM: case OP_U32_STORE_XOR_FP_IMM: case OP_U32_STORE_EQ_FP_IMM: return 5;
        case OP_U32_STORE_DIV_REM_FP_IMM: return 6;
        default: return 4;
    }
}

// ------------------------------------------------------------------ adapter
class MemoryModel {  // adapter::memory::Memory with dense address-indexed storage
   public:
    struct Cell {
        Word4 value;
        u32 clock, multiplicity;
        bool present = false;
    };
    std::vector<Cell> initial, final_;
    std::vector<ClockUpdateRow> clock_update_data;

    explicit MemoryModel(const std::vector<Word4>& initial_memory) {
        initial.resize(initial_memory.size());
        for (size_t a = 0; a < initial_memory.size(); a++) {
            initial[a].value = initial_memory[a];
            initial[a].clock = 0;
            initial[a].multiplicity = 0;
            initial[a].present = true;
        }
        final_ = initial;
    }
    struct Arg {
        u32 address, prev_clock, clock;
        Word4 prev_val, value;
    };
    Arg push(u32 address, Word4 value, u32 clock) {
        if (address >= (1u << 28)) throw std::runtime_error("adapter: memory address outside the 2^28-cell address space");  // TREE_HEIGHT 30, 4 words per cell
        if (address >= final_.size()) {
            final_.resize(address + 1);
            initial.resize(address + 1);
        }
        Cell prev;
        if (final_[address].present) prev = final_[address];
        else {
            prev.value = value;
            prev.clock = 0;
            prev.multiplicity = P - 1;
        }
        final_[address].value = value;
        final_[address].clock = clock;
        final_[address].multiplicity = P - 1;
        final_[address].present = true;
        u32 prev_clk = prev.clock;
        if (prev_clk == 0) {
            if (initial[address].present) initial[address].multiplicity = 1;
            else {
                initial[address].value = value;
                initial[address].clock = 0;
                initial[address].multiplicity = 1;
                initial[address].present = true;
            }
        }
        if (clock > prev_clk) {
            u32 delta = clock - prev_clk;
            if (delta > RC20_LIMIT) {
                u32 num_steps = delta / RC20_LIMIT;
                for (u32 s = 0; s < num_steps; s++) {
                    ClockUpdateRow r;
                    r.address = address;
                    r.prev_clk = prev_clk;
                    for (int k = 0; k < 4; k++) r.value[k] = initial[address].value.v[k];
                    clock_update_data.push_back(r);
                    prev_clk = m31_add(prev_clk, RC20_LIMIT);
                }
            }
        }
        return Arg{address, prev_clk, clock, prev.value, value};
    }
    void update_multiplicities(const PublicRanges& r) {
        auto fix_in = [&](u32 a) {
            if (a < initial.size() && initial[a].present) initial[a].multiplicity = 0;
            if (a < final_.size() && final_[a].present && final_[a].multiplicity == 0) final_[a].multiplicity = P - 1;
        };
        for (u32 a = r.program_start; a < r.program_end; a++) fix_in(a);
        for (u32 a = r.input_start; a < r.input_end; a++) fix_in(a);
        for (u32 a = r.output_start; a < r.output_end; a++) {
            if (a < final_.size() && final_[a].present) final_[a].multiplicity = 0;
            if (a < initial.size() && initial[a].present) initial[a].multiplicity = 1;
        }
    }
};

// build_partial_merkle_tree (crates/prover/src/adapter/merkle.rs:163-258): leaves = the 4 M31 words of every present
// cell at depth 30 (multiplicity 2 inside the public ranges of the tree's side, else 1); per depth 30..1, pairs in ascending
// index order, a missing sibling is the default hash of that depth with multiplicity 0; parents get multiplicity 1.
inline const std::vector<u32>& poseidon2_default_hashes() {  // Poseidon2Hash::default_hashes (poseidon2.rs:35-56)
    static const std::vector<u32> d = [] {
        std::vector<u32> v(TREE_HEIGHT + 1, 0);
        for (int depth = (int)TREE_HEIGHT - 1; depth >= 0; depth--) v[depth] = poseidon2_hash(v[depth + 1], v[depth + 1]);
        return v;
    }();
    return d;
}
struct MerkleLeafCell {
    u32 address;
    u32 value[4];
};
inline u32 build_partial_merkle_tree(const std::vector<MerkleLeafCell>& cells, bool initial_tree, const PublicRanges& r,
                                     std::vector<MerkleNode>& nodes_out) {
    if (cells.empty()) throw std::runtime_error("adapter: empty memory has no Merkle root");
    struct Val {
        u32 value, multiplicity;
    };
    std::map<u32, Val> cur;  // ordered: the reference sorts the indices of every depth
    for (const MerkleLeafCell& c : cells) {
        if (c.address >= (1u << 28)) throw std::runtime_error("adapter: address outside the 2^28-cell memory");
        bool is_public = initial_tree ? ((c.address >= r.program_start && c.address < r.program_end) || (c.address >= r.input_start && c.address < r.input_end))
                                      : (c.address >= r.output_start && c.address < r.output_end);
        for (u32 k = 0; k < 4; k++) cur[(c.address << 2) + k] = Val{c.value[k], is_public ? 2u : 1u};
    }
    const std::vector<u32>& defaults = poseidon2_default_hashes();
    size_t first = nodes_out.size();
    for (u32 depth = TREE_HEIGHT; depth >= 1; depth--) {
        std::map<u32, Val> parents;
        for (auto it = cur.begin(); it != cur.end(); ++it) {
            u32 index = it->first, left_index = index & ~1u, right_index = left_index | 1u;
            if (parents.count(index >> 1)) continue;  // the pair was processed with its left sibling
            auto l = cur.find(left_index), rr = cur.find(right_index);
            Val left = l != cur.end() ? l->second : Val{defaults[depth], 0u};
            Val right = rr != cur.end() ? rr->second : Val{defaults[depth], 0u};
            Val parent{poseidon2_hash(left.value, right.value), 1u};
            nodes_out.push_back(MerkleNode{left_index, depth, left.value, right.value, parent.value, left.multiplicity, right.multiplicity,
                                           parent.multiplicity, 0u});
            parents[index >> 1] = parent;
        }
        cur.swap(parents);
    }
    if (cur.size() != 1) throw std::logic_error("adapter: Merkle tree did not reduce to one root");
    u32 root = cur.begin()->second.value;
    for (size_t i = first; i < nodes_out.size(); i++) nodes_out[i].root = root;
    return root;
}

// Tail of import_internal (adapter/mod.rs:134-174): boundary multiplicities, the partial Poseidon2 Merkle trees of the
// initial and the final memory, the boundary rows in ascending address order.  Shared by the host adapter below and the
// device adapter (csrc/adapter.cu), which fills `memory` from the distinct cells it resolved on the GPU.
inline void finish_memory(MemoryModel& memory, const PublicRanges& ranges, ProverInput& in) {
    memory.update_multiplicities(ranges);
    in.public_ranges = ranges;
    auto dump = [&](const std::vector<MemoryModel::Cell>& cells, u32 root, std::vector<MemoryRow>& out) {
        for (size_t a = 0; a < cells.size(); a++) {
            if (!cells[a].present) continue;
            MemoryRow r;
            r.address = (u32)a;
            r.clock = cells[a].clock;
            for (int k = 0; k < 4; k++) r.value[k] = cells[a].value.v[k];
            r.multiplicity = cells[a].multiplicity;
            r.root = root;
            out.push_back(r);
        }
    };
    // memory commitments (adapter/mod.rs:152-174): partial Poseidon2 Merkle trees of the initial and the final memory
    auto leaves = [&](const std::vector<MemoryModel::Cell>& cells) {
        std::vector<MerkleLeafCell> v;
        for (size_t a = 0; a < cells.size(); a++)
            if (cells[a].present) v.push_back(MerkleLeafCell{(u32)a, {cells[a].value.v[0], cells[a].value.v[1], cells[a].value.v[2], cells[a].value.v[3]}});
        return v;
    };
    in.initial_root = build_partial_merkle_tree(leaves(memory.initial), true, ranges, in.merkle_nodes);
    in.final_root = build_partial_merkle_tree(leaves(memory.final_), false, ranges, in.merkle_nodes);
    dump(memory.initial, in.initial_root, in.initial_memory);
    dump(memory.final_, in.final_root, in.final_memory);
}

inline ProverInput import_from_vm(const VmTrace& vm) {
    ProverInput in;
    if (vm.trace.size() < 2) throw std::runtime_error("adapter: empty trace");
    MemoryModel memory(vm.initial_memory);
    in.initial_registers = vm.trace.front();
    in.final_registers = vm.trace.back();
    size_t mi = 0;
    u32 clock = 1;  // clock 0 is reserved for preloaded values
    auto next_mem = [&]() -> const std::pair<u32, Word4>& {
        if (mi >= vm.memory_trace.size()) throw std::runtime_error("adapter: unexpected end of the memory trace");
        return vm.memory_trace[mi++];
    };
    for (size_t s = 0; s + 1 < vm.trace.size(); s++) {
        const auto& ie = next_mem();
        MemoryModel::Arg iarg = memory.push(ie.first, ie.second, clock);
        u32 opcode = ie.second.v[0];
        int n_acc = opcode_memory_accesses(opcode);
        if (n_acc < 0) throw std::runtime_error("adapter: invalid opcode");
        int size_m31 = opcode_size_in_m31s(opcode);
        Bundle b;
        b.pc = vm.trace[s].pc;
        b.fp = vm.trace[s].fp;
        b.clock = clock;
        b.inst_prev_clock = iarg.prev_clock;
        for (int k = 0; k < 6; k++) b.inst[k] = (k < 4 && k < size_m31) ? ie.second.v[k] : 0;
        if (size_m31 > 4) {  // second QM31 word of the instruction, pushed at the same clock (adapter/memory.rs:317-339)
            const auto& ie2 = next_mem();
            memory.push(ie2.first, ie2.second, clock);
            b.inst[4] = ie2.second.v[0];
            if (size_m31 > 5) b.inst[5] = ie2.second.v[1];
        }
        b.span_start = (u32)in.data_accesses.size();
        for (int k = 0; k < n_acc; k++) {
            const auto& me = next_mem();
            MemoryModel::Arg a = memory.push(me.first, me.second, clock);
            in.data_accesses.push_back(DataAccess{a.address, a.prev_clock, a.prev_val.v[0], a.value.v[0]});
        }
        b.span_len = (u32)in.data_accesses.size() - b.span_start;
        in.states_by_opcodes[opcode].push_back(b);
        clock += 1;
    }
    in.n_steps = vm.trace.size() - 1;
    in.clock_update_data = memory.clock_update_data;
    finish_memory(memory, vm.public_ranges, in);
    return in;
}

}  // namespace cm31
