// cairo-m AIR components needed by `fibonacci_loop` (Cpu opcodes + Memory + ClockUpdate + RangeCheck),
// restated from the CODE of the reference, each as
//   evaluate<E>(E&)     — the FrameworkEval::evaluate body (constraints + relation entries)
//   write_trace<T>(T&)  — the witness row as expressions over the unpacked ExecutionBundle columns
// Column order = order of next_trace_mask() calls; constraint order = order of add_constraint
// calls followed by the logup constraints emitted by finalize_logup_in_pairs.
//
//   store_imm      crates/prover/src/components/opcodes/store_imm.rs:114-210 (write_trace), :325-419 (evaluate)
//   store_fp_imm   .../opcodes/store_fp_imm.rs:147-300, :457-619
//   store_fp_fp    .../opcodes/store_fp_fp.rs:153-309, :496-681
//   jnz_fp_imm     .../opcodes/jnz_fp_imm.rs:121-232, :351-446
//   jmp_imm        .../opcodes/jmp_imm.rs:110-190, :285-345
//   ret            .../opcodes/ret.rs:117-225, :386-486
//   assert_eq_fp_imm     .../opcodes/assert_eq_fp_imm.rs:150-200, :323-417
//   call_abs_imm         .../opcodes/call_abs_imm.rs:119-232, :381-494
//   store_frame_pointer  .../opcodes/store_frame_pointer.rs:147-198, :321-417
//   double_deref_fp_imm  .../opcodes/double_deref_fp_imm.rs:150-260, :395-509
//   double_deref_fp_fp   .../opcodes/double_deref_fp_fp.rs:175-275, :430-562
//   store_le_fp_imm      .../opcodes/store_le_fp_imm.rs:195-400, :560-747
//   u32_store_imm        .../opcodes/u32_store_imm.rs:160-215, :435-566
//   u32_store_add_fp_fp  .../opcodes/u32_store_add_fp_fp.rs:200-290, :590-809
//   u32_store_sub_fp_fp  .../opcodes/u32_store_sub_fp_fp.rs:200-290, :540-814
//   u32_store_lt_fp_fp   .../opcodes/u32_store_lt_fp_fp.rs:195-290, :560-794
//   u32_store_bitwise_fp_fp .../opcodes/u32_store_bitwise_fp_fp.rs:190-360, :560-783
//   u32_store_add_fp_imm, u32_store_lt_fp_imm, u32_store_eq_fp_fp, u32_store_eq_fp_imm, u32_store_mul_fp_fp,
//   u32_store_mul_fp_imm, u32_store_div_fp_fp, u32_store_div_fp_imm, u32_store_bitwise_fp_imm: cited at each struct
//   bitwise (table)      crates/prover/src/preprocessed/bitwise.rs:72-140 (multiplicities), :196-215 (evaluate), :253-290 (columns)
//   memory         crates/prover/src/components/memory.rs:93-195, :294-366
//   merkle         crates/prover/src/components/merkle.rs:73-170, :285-377
//   poseidon2      crates/prover/src/components/poseidon2.rs:143-290, :385-505 (constants: csrc/cairo/poseidon2_constants.hpp)
//   clock_update   crates/prover/src/components/clock_update.rs:70-160, :217-262
//   range_check_N  crates/prover/src/preprocessed/range_check/range_check_macro.rs:62-112, :171-183
#pragma once
#include <string>
#include <vector>

#include "../cairo/poseidon2.hpp"
#include "../field.cuh"

namespace cm31 {

// relation ids, in Relations::draw order (crates/prover/src/components/mod.rs:311-323)
enum CairoRelation : int { REL_REGISTERS = 0, REL_MEMORY, REL_MERKLE, REL_POSEIDON2, REL_RC8, REL_RC16, REL_RC20, REL_BITWISE, N_CAIRO_RELATIONS };
// relation sizes (crates/prover/src/relations.rs:7-44)
inline size_t cairo_relation_size(int r) {
    static const size_t sizes[N_CAIRO_RELATIONS] = {3, 6, 4, 16, 1, 1, 1, 4};
    return sizes[r];
}

// opcode ids (crates/common/src/instruction.rs:317-432)
constexpr u32 OP_STORE_ADD_FP_FP = 0, OP_STORE_SUB_FP_FP = 1, OP_STORE_MUL_FP_FP = 2, OP_STORE_DIV_FP_FP = 3;
constexpr u32 OP_STORE_ADD_FP_IMM = 4, OP_STORE_MUL_FP_IMM = 6, OP_STORE_IMM = 9, OP_CALL_ABS_IMM = 10, OP_RET = 11;
constexpr u32 OP_JMP_ABS_IMM = 12, OP_JMP_REL_IMM = 13, OP_JNZ_FP_IMM = 14;
constexpr u32 OP_STORE_DOUBLE_DEREF_FP = 8, OP_STORE_DOUBLE_DEREF_FP_FP = 42, OP_STORE_FRAME_POINTER = 43;
constexpr u32 OP_STORE_LE_FP_IMM = 48;
constexpr u32 OP_U32_STORE_ADD_FP_FP = 15, OP_U32_STORE_SUB_FP_FP = 16, OP_U32_STORE_IMM = 23;
constexpr u32 OP_U32_STORE_LT_FP_FP = 28;
constexpr u32 OP_U32_STORE_MUL_FP_FP = 17, OP_U32_STORE_DIV_REM_FP_FP = 18, OP_U32_STORE_ADD_FP_IMM = 19;
constexpr u32 OP_U32_STORE_MUL_FP_IMM = 21, OP_U32_STORE_DIV_REM_FP_IMM = 22, OP_U32_STORE_EQ_FP_FP = 24;
constexpr u32 OP_U32_STORE_EQ_FP_IMM = 30, OP_U32_STORE_LT_FP_IMM = 34;
constexpr u32 OP_U32_STORE_AND_FP_IMM = 39, OP_U32_STORE_OR_FP_IMM = 40, OP_U32_STORE_XOR_FP_IMM = 41;
constexpr u32 OP_U32_STORE_AND_FP_FP = 36, OP_U32_STORE_OR_FP_FP = 37, OP_U32_STORE_XOR_FP_FP = 38;

// Lookup tables (preprocessed columns with a multiplicity component): the table row a looked-up
// tuple lands on.  RangeCheckN: the value itself (range_check_macro.rs:72-84); Bitwise: the stacked
// index op * 2^16 + input1 * 2^8 + input2 (preprocessed/bitwise.rs:86-99).
constexpr u32 BITWISE_STACKED_LOG_SIZE = 18;
inline std::vector<u32> cairo_table_index_weights(int relation) {
    if (relation == REL_BITWISE) return {1u << 16, 1u << 8, 1u, 0u};
    return {1u};
}
constexpr u32 OP_STORE_TO_DOUBLE_DEREF_FP_IMM = 44, OP_STORE_TO_DOUBLE_DEREF_FP_FP = 45, OP_ASSERT_EQ_FP_IMM = 50;

constexpr u32 TREE_HEIGHT = 30;  // crates/prover/src/adapter/merkle.rs (memory address space 2^30)
constexpr u32 LOG_SIZE_RC_20 = 20;
constexpr u32 RC20_LIMIT = (1u << LOG_SIZE_RC_20) - 1;  // crates/prover/src/adapter/memory.rs:16

// unpacked ExecutionBundle columns (crates/prover/src/utils/execution_bundle.rs:12-75,
// crates/prover/src/utils/data_accesses.rs:10-28)
constexpr int IN_PC = 0, IN_FP = 1, IN_CLOCK = 2, IN_INST_PREV_CLOCK = 3, IN_INST0 = 4;
constexpr int IN_ACC_BASE = 10, ACC_ADDRESS = 0, ACC_PREV_CLOCK = 1, ACC_PREV_VALUE = 2, ACC_VALUE = 3;
constexpr int MAX_ACCESSES = 8;  // u32 DivRem touches 4 operands x 2 limbs
constexpr int N_BUNDLE_INPUTS = IN_ACC_BASE + 4 * MAX_ACCESSES;
inline int in_acc(int k, int field) { return IN_ACC_BASE + 4 * k + field; }

struct OpcodeEvalBase {
    u32 log_size_;
    u32 log_size() const { return log_size_; }
    u32 max_constraint_log_degree_bound() const { return log_size_ + 1; }
};

// ------------------------------------------------------------------ store_imm
struct StoreImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 9;
    static constexpr int N_ACCESSES = 1;
    static const char* name() { return "store_imm"; }
    static std::vector<u32> opcodes() { return {OP_STORE_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto zero = eval.f_const(0);
        auto opcode_constant = eval.f_const(OP_STORE_IMM);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto off0 = eval.next_trace_mask();
        auto off2 = eval.next_trace_mask();
        auto dst_prev_clock = eval.next_trace_mask();
        auto dst_prev_val = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, off0, off2});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, off0, off2});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + off2, dst_prev_clock, dst_prev_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + off2, clock, off0, zero, zero, zero});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        t.out(0, t.enabler());
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, t.in(IN_INST0 + 2));
        t.out(7, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(8, t.in(in_acc(0, ACC_PREV_VALUE)));
    }
};

// ------------------------------------------------------------------ store_fp_imm (StoreAddFpImm / StoreMulFpImm)
struct StoreFpImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 18;
    static constexpr int N_ACCESSES = 2;
    static const char* name() { return "store_fp_imm"; }
    static std::vector<u32> opcodes() { return {OP_STORE_ADD_FP_IMM, OP_STORE_MUL_FP_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto src_off = eval.next_trace_mask();
        auto imm = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto src_prev_clock = eval.next_trace_mask();
        auto src_val = eval.next_trace_mask();
        auto imm_inv = eval.next_trace_mask();
        auto dst_prev_clock = eval.next_trace_mask();
        auto dst_prev_val = eval.next_trace_mask();
        auto dst_val = eval.next_trace_mask();
        auto opcode_flag_0 = eval.next_trace_mask();
        auto opcode_flag_1 = eval.next_trace_mask();
        auto prod = eval.next_trace_mask();
        auto div = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(opcode_flag_0 * (one - opcode_flag_0));
        eval.add_constraint(opcode_flag_1 * (one - opcode_flag_1));
        eval.add_constraint(prod - src_val * imm);
        eval.add_constraint(imm * (imm_inv * imm - one));
        eval.add_constraint(imm_inv * (imm_inv * imm - one));
        eval.add_constraint(div - src_val * imm_inv);
        auto is_add = eval.add_intermediate((one - opcode_flag_0) * (one - opcode_flag_1));
        auto is_sub = eval.add_intermediate((one - opcode_flag_0) * opcode_flag_1);
        auto is_mul = eval.add_intermediate(opcode_flag_0 * (one - opcode_flag_1));
        auto is_div = eval.add_intermediate(opcode_flag_0 * opcode_flag_1);
        auto opcode_id = eval.add_intermediate(eval.f_const(OP_STORE_ADD_FP_IMM) + eval.f_const(2) * opcode_flag_0 + opcode_flag_1);
        auto res = eval.add_intermediate(is_add * (src_val + imm) + is_sub * (src_val - imm) + is_mul * prod + is_div * div);
        eval.add_constraint(dst_val - res);
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_id, src_off, imm, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_id, src_off, imm, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src_off, src_prev_clock, src_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src_off, clock, src_val});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock, dst_prev_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, dst_val});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - src_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto enabler = t.enabler();
        // padding rows carry the default bundle (a Ret with zero fields): imm = 0, flags * enabler = 0
        auto imm = t.in(IN_INST0 + 2);
        auto imm_inv = t.f_inv(imm);
        auto src_val = t.in(in_acc(0, ACC_VALUE));
        // flag = opcode.saturating_sub(STORE_ADD_FP_IMM); flag/2, flag%2, then * enabler
        auto flag = enabler * (t.in(IN_INST0) - t.f_const(OP_STORE_ADD_FP_IMM));
        t.out(0, enabler);
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, imm);
        t.out(7, t.in(IN_INST0 + 3));
        t.out(8, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(9, src_val);
        t.out(10, imm_inv);
        t.out(11, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(12, t.in(in_acc(1, ACC_PREV_VALUE)));
        t.out(13, t.in(in_acc(1, ACC_VALUE)));
        t.out(14, t.f_shr(flag, 1));
        t.out(15, t.f_and(flag, 1));
        t.out(16, src_val * imm);
        t.out(17, src_val * imm_inv);
    }
};

// ------------------------------------------------------------------ store_fp_fp (Add/Sub/Mul/Div)
struct StoreFpFpEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 20;
    static constexpr int N_ACCESSES = 3;
    static const char* name() { return "store_fp_fp"; }
    static std::vector<u32> opcodes() { return {OP_STORE_ADD_FP_FP, OP_STORE_SUB_FP_FP, OP_STORE_MUL_FP_FP, OP_STORE_DIV_FP_FP}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto off0 = eval.next_trace_mask();
        auto off1 = eval.next_trace_mask();
        auto off2 = eval.next_trace_mask();
        auto op0_prev_clock = eval.next_trace_mask();
        auto op0_val = eval.next_trace_mask();
        auto op1_prev_clock = eval.next_trace_mask();
        auto op1_val = eval.next_trace_mask();
        auto op1_inv = eval.next_trace_mask();
        auto dst_prev_clock = eval.next_trace_mask();
        auto dst_prev_val = eval.next_trace_mask();
        auto dst_val = eval.next_trace_mask();
        auto opcode_flag_0 = eval.next_trace_mask();
        auto opcode_flag_1 = eval.next_trace_mask();
        auto prod = eval.next_trace_mask();
        auto div = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(opcode_flag_0 * (one - opcode_flag_0));
        eval.add_constraint(opcode_flag_1 * (one - opcode_flag_1));
        eval.add_constraint(prod - op0_val * op1_val);
        eval.add_constraint(op1_val * (op1_inv * op1_val - one));
        eval.add_constraint(op1_inv * (op1_inv * op1_val - one));
        eval.add_constraint(div - op0_val * op1_inv);
        auto is_add = eval.add_intermediate((one - opcode_flag_0) * (one - opcode_flag_1));
        auto is_sub = eval.add_intermediate((one - opcode_flag_0) * opcode_flag_1);
        auto is_mul = eval.add_intermediate(opcode_flag_0 * (one - opcode_flag_1));
        auto is_div = eval.add_intermediate(opcode_flag_0 * opcode_flag_1);
        auto opcode_id = eval.add_intermediate(eval.f_const(OP_STORE_ADD_FP_FP) + eval.f_const(2) * opcode_flag_0 + opcode_flag_1);
        auto res = eval.add_intermediate(is_add * (op0_val + op1_val) + is_sub * (op0_val - op1_val) + is_mul * prod + is_div * div);
        eval.add_constraint(dst_val - res);
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_id, off0, off1, off2});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_id, off0, off1, off2});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + off0, op0_prev_clock, op0_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + off0, clock, op0_val});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + off1, op1_prev_clock, op1_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + off1, clock, op1_val});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + off2, dst_prev_clock, dst_prev_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + off2, clock, dst_val});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op1_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto enabler = t.enabler();
        auto op0_val = t.in(in_acc(0, ACC_VALUE));
        auto op1_val = t.in(in_acc(1, ACC_VALUE));
        auto op1_inv = t.f_inv(op1_val);
        // flag = (opcode == RET) ? 0 : opcode - STORE_ADD_FP_FP; only padding rows hold RET
        auto flag = enabler * (t.in(IN_INST0) - t.f_const(OP_STORE_ADD_FP_FP));
        t.out(0, enabler);
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, t.in(IN_INST0 + 2));
        t.out(7, t.in(IN_INST0 + 3));
        t.out(8, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(9, op0_val);
        t.out(10, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(11, op1_val);
        t.out(12, op1_inv);
        t.out(13, t.in(in_acc(2, ACC_PREV_CLOCK)));
        t.out(14, t.in(in_acc(2, ACC_PREV_VALUE)));
        t.out(15, t.in(in_acc(2, ACC_VALUE)));
        t.out(16, t.f_shr(flag, 1));
        t.out(17, t.f_and(flag, 1));
        t.out(18, op0_val * op1_val);
        t.out(19, op0_val * op1_inv);
    }
};

// ------------------------------------------------------------------ jnz_fp_imm
struct JnzFpImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 12;
    static constexpr int N_ACCESSES = 1;
    static const char* name() { return "jnz_fp_imm"; }
    static std::vector<u32> opcodes() { return {OP_JNZ_FP_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto opcode_constant = eval.f_const(OP_JNZ_FP_IMM);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto off0 = eval.next_trace_mask();
        auto imm = eval.next_trace_mask();
        auto op0_prev_clock = eval.next_trace_mask();
        auto op0_val = eval.next_trace_mask();
        auto op0_val_inv = eval.next_trace_mask();
        auto taken = eval.next_trace_mask();
        auto pc_new = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(enabler * op0_val * (taken - one));
        eval.add_constraint(enabler * (taken - op0_val * op0_val_inv));
        eval.add_constraint(enabler * (pc_new - pc - one - taken * (imm - one)));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc_new, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, off0, imm});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, off0, imm});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + off0, op0_prev_clock, op0_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + off0, clock, op0_val});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto one = t.f_const(1);
        auto pc = t.in(IN_PC);
        auto imm = t.in(IN_INST0 + 2);
        auto op0_val = t.in(in_acc(0, ACC_VALUE));
        auto op0_val_inv = t.f_inv(op0_val);
        auto taken = op0_val * op0_val_inv;  // 1 if op0 != 0 else 0
        t.out(0, t.enabler());
        t.out(1, pc);
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, imm);
        t.out(7, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(8, op0_val);
        t.out(9, op0_val_inv);
        t.out(10, taken);
        t.out(11, pc + one + taken * (imm - one));
    }
};

// ------------------------------------------------------------------ jmp_imm (JmpAbsImm / JmpRelImm)
struct JmpImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 7;
    static constexpr int N_ACCESSES = 0;
    static const char* name() { return "jmp_imm"; }
    static std::vector<u32> opcodes() { return {OP_JMP_ABS_IMM, OP_JMP_REL_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto opcode_constant = eval.f_const(OP_JMP_ABS_IMM);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto off0 = eval.next_trace_mask();
        auto is_rel = eval.next_trace_mask();
        auto opcode_id = opcode_constant + is_rel;
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(is_rel * (one - is_rel));
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_id, off0});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_id, off0});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {off0 + pc * is_rel, fp, clock + one});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto enabler = t.enabler();
        t.out(0, enabler);
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, enabler * (t.in(IN_INST0) - t.f_const(OP_JMP_ABS_IMM)));
    }
};

// ------------------------------------------------------------------ ret
struct RetEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 9;
    static constexpr int N_ACCESSES = 2;
    static const char* name() { return "ret"; }
    static std::vector<u32> opcodes() { return {OP_RET}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto two = eval.f_const(2);
        auto opcode_constant = eval.f_const(OP_RET);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto fp_min_2_prev_clock = eval.next_trace_mask();
        auto fp_min_2_val = eval.next_trace_mask();
        auto fp_min_1_prev_clock = eval.next_trace_mask();
        auto fp_min_1_val = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {fp_min_1_val, fp_min_2_val, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp - two, fp_min_2_prev_clock, fp_min_2_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp - two, clock, fp_min_2_val});
        // the reference subtracts `enabler` (not the constant 1) here: ret.rs:452-462
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp - enabler, fp_min_1_prev_clock, fp_min_1_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp - enabler, clock, fp_min_1_val});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - fp_min_2_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - fp_min_1_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        // the VM reads [fp-1] first, then [fp-2] (crates/runner/src/vm/instructions/call.rs:69-78)
        t.out(0, t.enabler());
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(6, t.in(in_acc(1, ACC_VALUE)));
        t.out(7, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(8, t.in(in_acc(0, ACC_VALUE)));
    }
};

// ------------------------------------------------------------------ assert_eq_fp_imm
struct AssertEqFpImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 9;
    static const char* name() { return "assert_eq_fp_imm"; }
    static std::vector<u32> opcodes() { return {OP_ASSERT_EQ_FP_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto opcode_constant = eval.f_const(OP_ASSERT_EQ_FP_IMM);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto src0_off = eval.next_trace_mask();
        auto imm = eval.next_trace_mask();
        auto op0_prev_clock = eval.next_trace_mask();
        auto op0_val = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(op0_val - imm);
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src0_off, imm});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src0_off, imm});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off, op0_prev_clock, op0_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off, clock, op0_val});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        t.out(0, t.enabler());
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, t.in(IN_INST0 + 2));
        t.out(7, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(8, t.in(in_acc(0, ACC_VALUE)));
    }
};

// ------------------------------------------------------------------ call_abs_imm
struct CallAbsImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 11;
    static const char* name() { return "call_abs_imm"; }
    static std::vector<u32> opcodes() { return {OP_CALL_ABS_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto opcode_constant = eval.f_const(OP_CALL_ABS_IMM);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto off0 = eval.next_trace_mask();
        auto off1 = eval.next_trace_mask();
        auto op0_prev_clock = eval.next_trace_mask();
        auto op0_prev_val = eval.next_trace_mask();
        auto op0_plus_one_prev_clock = eval.next_trace_mask();
        auto op0_plus_one_prev_val = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {off1, fp + off0 + one + one, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, off0, off1});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, off0, off1});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + off0, op0_prev_clock, op0_prev_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + off0, clock, fp});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + off0 + one, op0_plus_one_prev_clock, op0_plus_one_prev_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + off0 + one, clock, pc + one});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_plus_one_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        t.out(0, t.enabler());
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, t.in(IN_INST0 + 2));
        t.out(7, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(8, t.in(in_acc(0, ACC_PREV_VALUE)));
        t.out(9, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(10, t.in(in_acc(1, ACC_PREV_VALUE)));
    }
};

// ------------------------------------------------------------------ store_frame_pointer
struct StoreFramePointerEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 9;
    static const char* name() { return "store_frame_pointer"; }
    static std::vector<u32> opcodes() { return {OP_STORE_FRAME_POINTER}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto opcode_constant = eval.f_const(OP_STORE_FRAME_POINTER);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto imm = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto dst_prev_val = eval.next_trace_mask();
        auto dst_prev_clock = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, imm, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, imm, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock, dst_prev_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, fp + imm});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        t.out(0, t.enabler());
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, t.in(IN_INST0 + 2));
        t.out(7, t.in(in_acc(0, ACC_PREV_VALUE)));
        t.out(8, t.in(in_acc(0, ACC_PREV_CLOCK)));
    }
};

// ------------------------------------------------------------------ double_deref_fp_imm
// StoreDoubleDerefFp: [fp+off2] = [[fp+off0]+off1];  StoreToDoubleDerefFpImm: [[fp+off0]+off1] = [fp+off2]
struct DoubleDerefFpImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 17;
    static const char* name() { return "double_deref_fp_imm"; }
    static std::vector<u32> opcodes() { return {OP_STORE_DOUBLE_DEREF_FP, OP_STORE_TO_DOUBLE_DEREF_FP_IMM}; }
    static u32 delta_inv() { return m31_inv(OP_STORE_TO_DOUBLE_DEREF_FP_IMM - OP_STORE_DOUBLE_DEREF_FP); }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto dinv = eval.f_const(delta_inv());
        auto base_opcode = eval.f_const(OP_STORE_DOUBLE_DEREF_FP);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto opcode_constant = eval.next_trace_mask();
        auto off0 = eval.next_trace_mask();
        auto off1 = eval.next_trace_mask();
        auto off2 = eval.next_trace_mask();
        auto val0 = eval.next_trace_mask();
        auto prev_clock0 = eval.next_trace_mask();
        auto addr1 = eval.next_trace_mask();
        auto val1 = eval.next_trace_mask();
        auto prev_clock1 = eval.next_trace_mask();
        auto addr2 = eval.next_trace_mask();
        auto prev_val2 = eval.next_trace_mask();
        auto prev_clock2 = eval.next_trace_mask();
        auto write_lhs = (opcode_constant - base_opcode) * dinv;
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(write_lhs * (one - write_lhs));
        eval.add_constraint(enabler * (addr1 - write_lhs * (fp + off2) - (one - write_lhs) * (val0 + off1)));
        eval.add_constraint(enabler * (addr2 - write_lhs * (val0 + off1) - (one - write_lhs) * (fp + off2)));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, off0, off1, off2});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, off0, off1, off2});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + off0, prev_clock0, val0});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + off0, clock, val0});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {addr1, prev_clock1, val1});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {addr1, clock, val1});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {addr2, prev_clock2, prev_val2});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {addr2, clock, val1});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - prev_clock0 - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - prev_clock1 - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - prev_clock2 - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto enabler = t.enabler();
        auto one = t.f_const(1);
        auto base_opcode = t.f_const(OP_STORE_DOUBLE_DEREF_FP);
        // padding rows hold the default bundle (a Ret): the reference rewrites their opcode to
        // STORE_DOUBLE_DEREF_FP (double_deref_fp_imm.rs:183-189)
        auto opcode_constant = enabler * (t.in(IN_INST0) - base_opcode) + base_opcode;
        auto fp = t.in(IN_FP);
        auto off1 = t.in(IN_INST0 + 2), off2 = t.in(IN_INST0 + 3);
        auto val0 = t.in(in_acc(0, ACC_VALUE));
        auto write_lhs = (opcode_constant - base_opcode) * t.f_const(delta_inv());
        auto addr1 = write_lhs * (fp + off2) + (one - write_lhs) * (val0 + off1);
        auto addr2 = write_lhs * (val0 + off1) + (one - write_lhs) * (fp + off2);
        t.out(0, enabler);
        t.out(1, t.in(IN_PC));
        t.out(2, fp);
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, opcode_constant);
        t.out(6, t.in(IN_INST0 + 1));
        t.out(7, off1);
        t.out(8, off2);
        t.out(9, val0);
        t.out(10, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(11, addr1);
        t.out(12, t.in(in_acc(1, ACC_VALUE)));
        t.out(13, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(14, addr2);
        t.out(15, t.in(in_acc(2, ACC_PREV_VALUE)));
        t.out(16, t.in(in_acc(2, ACC_PREV_CLOCK)));
    }
};

// ------------------------------------------------------------------ double_deref_fp_fp
// StoreDoubleDerefFpFp: [fp+off2] = [[fp+off0]+[fp+off1]];  StoreToDoubleDerefFpFp: the reverse
struct DoubleDerefFpFpEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 19;
    static const char* name() { return "double_deref_fp_fp"; }
    static std::vector<u32> opcodes() { return {OP_STORE_DOUBLE_DEREF_FP_FP, OP_STORE_TO_DOUBLE_DEREF_FP_FP}; }
    static u32 delta_inv() { return m31_inv(OP_STORE_TO_DOUBLE_DEREF_FP_FP - OP_STORE_DOUBLE_DEREF_FP_FP); }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto base_opcode = eval.f_const(OP_STORE_DOUBLE_DEREF_FP_FP);
        auto dinv = eval.f_const(delta_inv());
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto opcode_constant = eval.next_trace_mask();
        auto off0 = eval.next_trace_mask();
        auto off1 = eval.next_trace_mask();
        auto off2 = eval.next_trace_mask();
        auto val0 = eval.next_trace_mask();
        auto prev_clock0 = eval.next_trace_mask();
        auto val1 = eval.next_trace_mask();
        auto prev_clock1 = eval.next_trace_mask();
        auto addr2 = eval.next_trace_mask();
        auto val2 = eval.next_trace_mask();
        auto prev_clock2 = eval.next_trace_mask();
        auto addr3 = eval.next_trace_mask();
        auto prev_val3 = eval.next_trace_mask();
        auto prev_clock3 = eval.next_trace_mask();
        auto write_lhs = (opcode_constant - base_opcode) * dinv;
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(write_lhs * (one - write_lhs));
        eval.add_constraint(enabler * (addr2 - write_lhs * (fp + off2) - (one - write_lhs) * (val0 + val1)));
        eval.add_constraint(enabler * (addr3 - write_lhs * (val0 + val1) - (one - write_lhs) * (fp + off2)));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, off0, off1, off2});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, off0, off1, off2});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + off0, prev_clock0, val0});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + off0, clock, val0});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + off1, prev_clock1, val1});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + off1, clock, val1});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {addr2, prev_clock2, val2});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {addr2, clock, val2});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {addr3, prev_clock3, prev_val3});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {addr3, clock, val2});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - prev_clock0 - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - prev_clock1 - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - prev_clock2 - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - prev_clock3 - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto enabler = t.enabler();
        auto one = t.f_const(1);
        auto base_opcode = t.f_const(OP_STORE_DOUBLE_DEREF_FP_FP);
        auto opcode_constant = enabler * (t.in(IN_INST0) - base_opcode) + base_opcode;  // double_deref_fp_fp.rs:189-196
        auto fp = t.in(IN_FP);
        auto off2 = t.in(IN_INST0 + 3);
        auto val0 = t.in(in_acc(0, ACC_VALUE)), val1 = t.in(in_acc(1, ACC_VALUE));
        auto write_lhs = (opcode_constant - base_opcode) * t.f_const(delta_inv());
        auto addr2 = write_lhs * (fp + off2) + (one - write_lhs) * (val0 + val1);
        auto addr3 = write_lhs * (val0 + val1) + (one - write_lhs) * (fp + off2);
        t.out(0, enabler);
        t.out(1, t.in(IN_PC));
        t.out(2, fp);
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, opcode_constant);
        t.out(6, t.in(IN_INST0 + 1));
        t.out(7, t.in(IN_INST0 + 2));
        t.out(8, off2);
        t.out(9, val0);
        t.out(10, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(11, val1);
        t.out(12, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(13, addr2);
        t.out(14, t.in(in_acc(2, ACC_VALUE)));
        t.out(15, t.in(in_acc(2, ACC_PREV_CLOCK)));
        t.out(16, addr3);
        t.out(17, t.in(in_acc(3, ACC_PREV_VALUE)));
        t.out(18, t.in(in_acc(3, ACC_PREV_CLOCK)));
    }
};

// ------------------------------------------------------------------ u32_store_imm
// u32([fp+dst_off], [fp+dst_off+1]) = (imm_lo, imm_hi): a u32 lives in two consecutive cells as 16-bit limbs
struct U32StoreImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 12;
    static const char* name() { return "u32_store_imm"; }
    static std::vector<u32> opcodes() { return {OP_U32_STORE_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto opcode_constant = eval.f_const(OP_U32_STORE_IMM);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto imm_lo = eval.next_trace_mask();
        auto imm_hi = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto dst_prev_val_lo = eval.next_trace_mask();
        auto dst_prev_val_hi = eval.next_trace_mask();
        auto dst_prev_clock_lo = eval.next_trace_mask();
        auto dst_prev_clock_hi = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, imm_lo, imm_hi, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, imm_lo, imm_hi, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock_lo, dst_prev_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, imm_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off + one, dst_prev_clock_hi, dst_prev_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off + one, clock, imm_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {imm_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {imm_hi});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_hi - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        t.out(0, t.enabler());
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, t.in(IN_INST0 + 2));
        t.out(7, t.in(IN_INST0 + 3));
        t.out(8, t.in(in_acc(0, ACC_PREV_VALUE)));
        t.out(9, t.in(in_acc(1, ACC_PREV_VALUE)));
        t.out(10, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(11, t.in(in_acc(1, ACC_PREV_CLOCK)));
    }
};

// ------------------------------------------------------------------ u32_store_lt_fp_fp
// [fp+dst_off] = u32(op0) < u32(op1): no borrow out of op1 - op0 - 1, limb differences range-checked
struct U32StoreLtFpFpEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 20;
    static const char* name() { return "u32_store_lt_fp_fp"; }
    static std::vector<u32> opcodes() { return {OP_U32_STORE_LT_FP_FP}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto two_pow_16 = eval.f_const(1u << 16);
        auto opcode_constant = eval.f_const(OP_U32_STORE_LT_FP_FP);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto src0_off = eval.next_trace_mask();
        auto src1_off = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto op0_val_lo = eval.next_trace_mask();
        auto op0_val_hi = eval.next_trace_mask();
        auto op0_prev_clock_lo = eval.next_trace_mask();
        auto op0_prev_clock_hi = eval.next_trace_mask();
        auto op1_val_lo = eval.next_trace_mask();
        auto op1_val_hi = eval.next_trace_mask();
        auto op1_prev_clock_lo = eval.next_trace_mask();
        auto op1_prev_clock_hi = eval.next_trace_mask();
        auto dst_prev_val = eval.next_trace_mask();
        auto dst_prev_clock = eval.next_trace_mask();
        auto borrow_lo = eval.next_trace_mask();
        auto borrow_hi = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(borrow_lo * (one - borrow_lo));
        eval.add_constraint(borrow_hi * (one - borrow_hi));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src0_off, src1_off, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src0_off, src1_off, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off, op0_prev_clock_lo, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off, clock, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off + one, op0_prev_clock_hi, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off + one, clock, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off, op1_prev_clock_lo, op1_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off, clock, op1_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off + one, op1_prev_clock_hi, op1_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off + one, clock, op1_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock, dst_prev_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, one - borrow_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op1_val_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op1_val_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op1_val_lo - enabler + borrow_lo * two_pow_16 - op0_val_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op1_val_hi - borrow_lo + borrow_hi * two_pow_16 - op0_val_hi});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op1_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op1_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto enabler = t.enabler();
        auto one = t.f_const(1);
        auto op0_lo = t.in(in_acc(0, ACC_VALUE)), op0_hi = t.in(in_acc(1, ACC_VALUE));
        auto op1_lo = t.in(in_acc(2, ACC_VALUE)), op1_hi = t.in(in_acc(3, ACC_VALUE));
        // borrows of op1 - op0 - enabler (u32_store_lt_fp_fp.rs:222-256): x < y + z
        auto borrow_lo = one - t.f_le(op0_lo + enabler, op1_lo);
        auto borrow_hi = one - t.f_le(op0_hi + borrow_lo, op1_hi);
        t.out(0, enabler);
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, t.in(IN_INST0 + 2));
        t.out(7, t.in(IN_INST0 + 3));
        t.out(8, op0_lo);
        t.out(9, op0_hi);
        t.out(10, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(11, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(12, op1_lo);
        t.out(13, op1_hi);
        t.out(14, t.in(in_acc(2, ACC_PREV_CLOCK)));
        t.out(15, t.in(in_acc(3, ACC_PREV_CLOCK)));
        t.out(16, t.in(in_acc(4, ACC_PREV_VALUE)));
        t.out(17, t.in(in_acc(4, ACC_PREV_CLOCK)));
        t.out(18, borrow_lo);
        t.out(19, borrow_hi);
    }
};

// ------------------------------------------------------------------ u32_store_add_fp_fp / u32_store_sub_fp_fp
// dst = op0 +/- op1 on 16-bit limbs with carry / borrow bits; limbs and results range-checked (RangeCheck16).
// Shared body: SUB selects the borrow form (u32_store_sub_fp_fp.rs:572-575) instead of the carry form
// (u32_store_add_fp_fp.rs res_lo/res_hi).
template <bool SUB>
struct U32StoreBinFpFpEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 22;
    static const char* name() { return SUB ? "u32_store_sub_fp_fp" : "u32_store_add_fp_fp"; }
    static std::vector<u32> opcodes() { return {SUB ? OP_U32_STORE_SUB_FP_FP : OP_U32_STORE_ADD_FP_FP}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto two_pow_16 = eval.f_const(1u << 16);
        auto opcode_constant = eval.f_const(SUB ? OP_U32_STORE_SUB_FP_FP : OP_U32_STORE_ADD_FP_FP);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto src0_off = eval.next_trace_mask();
        auto src1_off = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto op0_val_lo = eval.next_trace_mask();
        auto op0_val_hi = eval.next_trace_mask();
        auto op0_prev_lo_clock = eval.next_trace_mask();
        auto op0_prev_hi_clock = eval.next_trace_mask();
        auto op1_val_lo = eval.next_trace_mask();
        auto op1_val_hi = eval.next_trace_mask();
        auto op1_prev_lo_clock = eval.next_trace_mask();
        auto op1_prev_hi_clock = eval.next_trace_mask();
        auto dst_prev_val_lo = eval.next_trace_mask();
        auto dst_prev_val_hi = eval.next_trace_mask();
        auto dst_prev_lo_clock = eval.next_trace_mask();
        auto dst_prev_hi_clock = eval.next_trace_mask();
        auto c_lo = eval.next_trace_mask();  // u16_carry | borrow_lo
        auto c_hi = eval.next_trace_mask();  // u32_carry | borrow_hi
        auto res_lo = SUB ? op0_val_lo + c_lo * two_pow_16 - op1_val_lo : op0_val_lo + op1_val_lo - c_lo * two_pow_16;
        auto res_hi = SUB ? op0_val_hi - c_lo + c_hi * two_pow_16 - op1_val_hi : op0_val_hi + op1_val_hi + c_lo - c_hi * two_pow_16;
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(c_lo * (one - c_lo));
        eval.add_constraint(c_hi * (one - c_hi));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src0_off, src1_off, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src0_off, src1_off, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off, op0_prev_lo_clock, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off, clock, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off + one, op0_prev_hi_clock, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off + one, clock, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off, op1_prev_lo_clock, op1_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off, clock, op1_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off + one, op1_prev_hi_clock, op1_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off + one, clock, op1_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_lo_clock, dst_prev_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, res_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off + one, dst_prev_hi_clock, dst_prev_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off + one, clock, res_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op1_val_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op1_val_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {res_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {res_hi});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_lo_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_hi_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op1_prev_lo_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op1_prev_hi_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_lo_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_hi_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto one = t.f_const(1);
        auto op0_lo = t.in(in_acc(0, ACC_VALUE)), op0_hi = t.in(in_acc(1, ACC_VALUE));
        auto op1_lo = t.in(in_acc(2, ACC_VALUE)), op1_hi = t.in(in_acc(3, ACC_VALUE));
        // carries: limb sum > 0xFFFF (u32_store_add_fp_fp.rs:226-243); borrows: x < y (u32_store_sub_fp_fp.rs:229-249)
        auto c_lo = SUB ? one - t.f_le(op1_lo, op0_lo) : t.f_le(t.f_const(1u << 16), op0_lo + op1_lo);
        auto c_hi = SUB ? one - t.f_le(op1_hi + c_lo, op0_hi) : t.f_le(t.f_const(1u << 16), op0_hi + op1_hi + c_lo);
        t.out(0, t.enabler());
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, t.in(IN_INST0 + 2));
        t.out(7, t.in(IN_INST0 + 3));
        t.out(8, op0_lo);
        t.out(9, op0_hi);
        t.out(10, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(11, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(12, op1_lo);
        t.out(13, op1_hi);
        t.out(14, t.in(in_acc(2, ACC_PREV_CLOCK)));
        t.out(15, t.in(in_acc(3, ACC_PREV_CLOCK)));
        t.out(16, t.in(in_acc(4, ACC_PREV_VALUE)));
        t.out(17, t.in(in_acc(5, ACC_PREV_VALUE)));
        t.out(18, t.in(in_acc(4, ACC_PREV_CLOCK)));
        t.out(19, t.in(in_acc(5, ACC_PREV_CLOCK)));
        t.out(20, c_lo);
        t.out(21, c_hi);
    }
};
typedef U32StoreBinFpFpEval<false> U32StoreAddFpFpEval;
typedef U32StoreBinFpFpEval<true> U32StoreSubFpFpEval;

// ------------------------------------------------------------------ u32_store_bitwise_fp_fp (And / Or / Xor)
// operands and result are decomposed into bytes; each byte triple is looked up in the Bitwise table
struct U32StoreBitwiseFpFpEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 29;
    static const char* name() { return "u32_store_bitwise_fp_fp"; }
    static std::vector<u32> opcodes() { return {OP_U32_STORE_AND_FP_FP, OP_U32_STORE_OR_FP_FP, OP_U32_STORE_XOR_FP_FP}; }
    template <class E>
    void evaluate(E& eval) const {
        auto two_pow_8 = eval.f_const(1u << 8);
        auto one = eval.f_const(1);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto opcode_constant = eval.next_trace_mask();
        auto src0_off = eval.next_trace_mask();
        auto src1_off = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto op0_val_0 = eval.next_trace_mask();
        auto op0_val_1 = eval.next_trace_mask();
        auto op0_val_2 = eval.next_trace_mask();
        auto op0_val_3 = eval.next_trace_mask();
        auto op0_prev_clock_lo = eval.next_trace_mask();
        auto op0_prev_clock_hi = eval.next_trace_mask();
        auto op1_val_0 = eval.next_trace_mask();
        auto op1_val_1 = eval.next_trace_mask();
        auto op1_val_2 = eval.next_trace_mask();
        auto op1_val_3 = eval.next_trace_mask();
        auto op1_prev_clock_lo = eval.next_trace_mask();
        auto op1_prev_clock_hi = eval.next_trace_mask();
        auto dst_prev_val_lo = eval.next_trace_mask();
        auto dst_prev_val_hi = eval.next_trace_mask();
        auto dst_val_0 = eval.next_trace_mask();
        auto dst_val_1 = eval.next_trace_mask();
        auto dst_val_2 = eval.next_trace_mask();
        auto dst_val_3 = eval.next_trace_mask();
        auto dst_prev_clock_lo = eval.next_trace_mask();
        auto dst_prev_clock_hi = eval.next_trace_mask();
        eval.add_constraint(enabler * (enabler - one));  // sign as in the reference (u32_store_bitwise_fp_fp.rs evaluate)
        auto bitwise_op = opcode_constant - eval.f_const(OP_U32_STORE_AND_FP_FP);
        auto op0_val_lo = op0_val_0 + op0_val_1 * two_pow_8;
        auto op0_val_hi = op0_val_2 + op0_val_3 * two_pow_8;
        auto op1_val_lo = op1_val_0 + op1_val_1 * two_pow_8;
        auto op1_val_hi = op1_val_2 + op1_val_3 * two_pow_8;
        auto dst_val_lo = dst_val_0 + dst_val_1 * two_pow_8;
        auto dst_val_hi = dst_val_2 + dst_val_3 * two_pow_8;
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src0_off, src1_off, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src0_off, src1_off, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off, op0_prev_clock_lo, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off, clock, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off + one, op0_prev_clock_hi, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off + one, clock, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off, op1_prev_clock_lo, op1_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off, clock, op1_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off + one, op1_prev_clock_hi, op1_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off + one, clock, op1_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock_lo, dst_prev_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, dst_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off + one, dst_prev_clock_hi, dst_prev_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off + one, clock, dst_val_hi});
        eval.add_to_relation(REL_BITWISE, -eval.ef_one(), {bitwise_op, op0_val_0, op1_val_0, dst_val_0});
        eval.add_to_relation(REL_BITWISE, -eval.ef_one(), {bitwise_op, op0_val_1, op1_val_1, dst_val_1});
        eval.add_to_relation(REL_BITWISE, -eval.ef_one(), {bitwise_op, op0_val_2, op1_val_2, dst_val_2});
        eval.add_to_relation(REL_BITWISE, -eval.ef_one(), {bitwise_op, op0_val_3, op1_val_3, dst_val_3});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op1_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op1_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_hi - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto enabler = t.enabler();
        auto base_opcode = t.f_const(OP_U32_STORE_AND_FP_FP);
        // padding rows (default bundle = Ret) are rewritten to U32_STORE_AND_FP_FP (u32_store_bitwise_fp_fp.rs:197-203)
        auto opcode_constant = enabler * (t.in(IN_INST0) - base_opcode) + base_opcode;
        auto lo8 = [&](decltype(enabler) v) { return t.f_and(v, 0xff); };
        auto hi8 = [&](decltype(enabler) v) { return t.f_and(t.f_shr(v, 8), 0xff); };
        auto op0_lo = t.in(in_acc(0, ACC_VALUE)), op0_hi = t.in(in_acc(1, ACC_VALUE));
        auto op1_lo = t.in(in_acc(2, ACC_VALUE)), op1_hi = t.in(in_acc(3, ACC_VALUE));
        auto dst_lo = t.in(in_acc(4, ACC_VALUE)), dst_hi = t.in(in_acc(5, ACC_VALUE));
        t.out(0, enabler);
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, opcode_constant);
        t.out(6, t.in(IN_INST0 + 1));
        t.out(7, t.in(IN_INST0 + 2));
        t.out(8, t.in(IN_INST0 + 3));
        t.out(9, lo8(op0_lo));
        t.out(10, hi8(op0_lo));
        t.out(11, lo8(op0_hi));
        t.out(12, hi8(op0_hi));
        t.out(13, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(14, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(15, lo8(op1_lo));
        t.out(16, hi8(op1_lo));
        t.out(17, lo8(op1_hi));
        t.out(18, hi8(op1_hi));
        t.out(19, t.in(in_acc(2, ACC_PREV_CLOCK)));
        t.out(20, t.in(in_acc(3, ACC_PREV_CLOCK)));
        t.out(21, t.in(in_acc(4, ACC_PREV_VALUE)));
        t.out(22, t.in(in_acc(5, ACC_PREV_VALUE)));
        t.out(23, lo8(dst_lo));
        t.out(24, hi8(dst_lo));
        t.out(25, lo8(dst_hi));
        t.out(26, hi8(dst_hi));
        t.out(27, t.in(in_acc(4, ACC_PREV_CLOCK)));
        t.out(28, t.in(in_acc(5, ACC_PREV_CLOCK)));
    }
};

// ------------------------------------------------------------------ store_le_fp_imm
// [fp+dst_off] = ([fp+src_off] <= imm), proven with the arc argument of cairo-lang's assert_le_felt
// (store_le_fp_imm.rs:1-95): of the three arcs a, b-a, P-1-b (a = min, b = max of the operands) the two
// shortest are range-checked as 16-bit limb pairs.
struct StoreLeFpImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 22;
    static constexpr u32 PRIME_OVER_3_HIGH = ((P / 3) >> 16) + 1;  // store_le_fp_imm.rs:132-133
    static constexpr u32 PRIME_OVER_2_HIGH = ((P / 2) >> 16) + 1;
    static const char* name() { return "store_le_fp_imm"; }
    static std::vector<u32> opcodes() { return {OP_STORE_LE_FP_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto opcode_constant = eval.f_const(OP_STORE_LE_FP_IMM);
        auto prime_over_3_high = eval.f_const(PRIME_OVER_3_HIGH);
        auto prime_over_2_high = eval.f_const(PRIME_OVER_2_HIGH);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto src_off = eval.next_trace_mask();
        auto imm = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto src_val = eval.next_trace_mask();
        auto src_prev_clock = eval.next_trace_mask();
        auto dst_prev_val = eval.next_trace_mask();
        auto dst_prev_clock = eval.next_trace_mask();
        auto a = eval.next_trace_mask();
        auto b = eval.next_trace_mask();
        auto keep_0_1 = eval.next_trace_mask();
        auto keep_0_2 = eval.next_trace_mask();
        auto keep_1_2 = eval.next_trace_mask();
        auto arc_short_lo = eval.next_trace_mask();
        auto arc_short_hi = eval.next_trace_mask();
        auto arc_long_lo = eval.next_trace_mask();
        auto arc_long_hi = eval.next_trace_mask();
        auto is_le = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(keep_0_1 * (one - keep_0_1));
        eval.add_constraint(keep_0_2 * (one - keep_0_2));
        eval.add_constraint(keep_1_2 * (one - keep_1_2));
        eval.add_constraint(enabler * (keep_0_1 + keep_0_2 + keep_1_2 - one));
        eval.add_constraint(is_le * (one - is_le));
        auto arc_short = arc_short_lo + arc_short_hi * prime_over_3_high;
        auto arc_long = arc_long_lo + arc_long_hi * prime_over_2_high;
        auto arc_sum = arc_short + arc_long;
        auto arc_prod = arc_short * arc_long;
        eval.add_constraint(keep_0_1 * (arc_sum - (a + b - a)));
        eval.add_constraint(keep_0_1 * (arc_prod - a * (b - a)));
        eval.add_constraint(keep_0_2 * (arc_sum - (a - one - b)));
        eval.add_constraint(keep_0_2 * (arc_prod - a * (-one - b)));
        eval.add_constraint(keep_1_2 * (arc_sum - (b - a - one - b)));
        eval.add_constraint(keep_1_2 * (arc_prod - (b - a) * (-one - b)));
        eval.add_constraint(enabler * (a - is_le * src_val - (one - is_le) * imm));
        eval.add_constraint(enabler * (b - is_le * imm - (one - is_le) * src_val));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src_off, imm, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src_off, imm, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src_off, src_prev_clock, src_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src_off, clock, src_val});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock, dst_prev_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, is_le});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {arc_short_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {arc_short_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {arc_long_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {arc_long_hi});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - src_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto enabler = t.enabler();
        auto one = t.f_const(1), two = t.f_const(2);
        auto half = t.f_const(m31_inv(2));
        auto src_val = t.in(in_acc(0, ACC_VALUE));
        auto imm = t.in(IN_INST0 + 2);
        auto is_le = t.f_le(src_val, imm);  // also 1 on padding rows (0 <= 0), as in the reference
        auto a = is_le * src_val + (one - is_le) * imm;  // min
        auto b = is_le * imm + (one - is_le) * src_val;  // max
        // the three arcs and their rank in the reference's STABLE ascending sort (sort_by_key on the
        // length, ties keep index order): rank_i = #arcs placed before arc i
        auto l0 = a, l1 = b - a, l2 = t.f_const(P - 1) - b;
        auto lt = [&](decltype(a) x, decltype(a) y) { return one - t.f_le(y, x); };
        auto r0 = lt(l1, l0) + lt(l2, l0);
        auto r1 = t.f_le(l0, l1) + lt(l2, l1);
        auto r2 = t.f_le(l0, l2) + t.f_le(l1, l2);
        auto is0 = [&](decltype(a) r) { return (r - one) * (r - two) * half; };
        auto is1 = [&](decltype(a) r) { return r * (two - r); };
        auto is2 = [&](decltype(a) r) { return r * (r - one) * half; };
        auto arc_short = is0(r0) * l0 + is0(r1) * l1 + is0(r2) * l2;
        auto arc_long = is1(r0) * l0 + is1(r1) * l1 + is1(r2) * l2;
        t.out(0, enabler);
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, imm);
        t.out(7, t.in(IN_INST0 + 3));
        t.out(8, src_val);
        t.out(9, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(10, t.in(in_acc(1, ACC_PREV_VALUE)));
        t.out(11, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(12, a);
        t.out(13, b);
        t.out(14, enabler * is2(r2));  // exclude == 2: keep arcs 0 and 1
        t.out(15, enabler * is2(r1));  // exclude == 1
        t.out(16, enabler * is2(r0));  // exclude == 0
        t.out(17, t.f_modc(arc_short, PRIME_OVER_3_HIGH));
        t.out(18, t.f_divc(arc_short, PRIME_OVER_3_HIGH));
        t.out(19, t.f_modc(arc_long, PRIME_OVER_2_HIGH));
        t.out(20, t.f_divc(arc_long, PRIME_OVER_2_HIGH));
        t.out(21, is_le);
    }
};

// ------------------------------------------------------------------ two-word u32 instructions
// The *_fp_imm u32 opcodes (and DivRem) carry 5-6 M31s: they span two QM31 words at pc, pc+1 (registers advance by
// 2), the second word is read with the same inst_prev_clock (e.g. u32_store_add_fp_imm.rs:568-586).

// ------------------------------------------------------------------ u32_store_add_fp_imm
//   .../opcodes/u32_store_add_fp_imm.rs:186-335 (write_trace), :520-742 (evaluate)
struct U32StoreAddFpImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 19;
    static const char* name() { return "u32_store_add_fp_imm"; }
    static std::vector<u32> opcodes() { return {OP_U32_STORE_ADD_FP_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto two_pow_16 = eval.f_const(1u << 16);
        auto opcode_constant = eval.f_const(OP_U32_STORE_ADD_FP_IMM);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto src_off = eval.next_trace_mask();
        auto imm_lo = eval.next_trace_mask();
        auto imm_hi = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto op0_val_lo = eval.next_trace_mask();
        auto op0_val_hi = eval.next_trace_mask();
        auto op0_prev_clock_lo = eval.next_trace_mask();
        auto op0_prev_clock_hi = eval.next_trace_mask();
        auto dst_prev_val_lo = eval.next_trace_mask();
        auto dst_prev_val_hi = eval.next_trace_mask();
        auto dst_prev_clock_lo = eval.next_trace_mask();
        auto dst_prev_clock_hi = eval.next_trace_mask();
        auto u16_carry = eval.next_trace_mask();
        auto u32_carry = eval.next_trace_mask();
        auto res_lo = op0_val_lo + imm_lo - u16_carry * two_pow_16;
        auto res_hi = op0_val_hi + imm_hi + u16_carry - u32_carry * two_pow_16;
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(u16_carry * (one - u16_carry));
        eval.add_constraint(u32_carry * (one - u32_carry));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src_off, imm_lo, imm_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src_off, imm_lo, imm_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc + one, inst_prev_clock, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc + one, clock, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src_off, op0_prev_clock_lo, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src_off, clock, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src_off + one, op0_prev_clock_hi, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src_off + one, clock, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock_lo, dst_prev_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, res_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off + one, dst_prev_clock_hi, dst_prev_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off + one, clock, res_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {imm_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {imm_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {res_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {res_hi});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_hi - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto one = t.f_const(1);
        auto limb_max = t.f_const(0xffff);
        auto imm_lo = t.in(IN_INST0 + 2), imm_hi = t.in(IN_INST0 + 3);
        auto op0_lo = t.in(in_acc(0, ACC_VALUE)), op0_hi = t.in(in_acc(1, ACC_VALUE));
        auto u16_carry = one - t.f_le(op0_lo + imm_lo, limb_max);
        auto u32_carry = one - t.f_le(op0_hi + imm_hi + u16_carry, limb_max);
        t.out(0, t.enabler());
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, imm_lo);
        t.out(7, imm_hi);
        t.out(8, t.in(IN_INST0 + 4));
        t.out(9, op0_lo);
        t.out(10, op0_hi);
        t.out(11, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(12, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(13, t.in(in_acc(2, ACC_PREV_VALUE)));
        t.out(14, t.in(in_acc(3, ACC_PREV_VALUE)));
        t.out(15, t.in(in_acc(2, ACC_PREV_CLOCK)));
        t.out(16, t.in(in_acc(3, ACC_PREV_CLOCK)));
        t.out(17, u16_carry);
        t.out(18, u32_carry);
    }
};

// ------------------------------------------------------------------ u32_store_lt_fp_imm
//   .../opcodes/u32_store_lt_fp_imm.rs:188-331, :492-694
struct U32StoreLtFpImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 17;
    static const char* name() { return "u32_store_lt_fp_imm"; }
    static std::vector<u32> opcodes() { return {OP_U32_STORE_LT_FP_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto two_pow_16 = eval.f_const(1u << 16);
        auto opcode_constant = eval.f_const(OP_U32_STORE_LT_FP_IMM);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto src_off = eval.next_trace_mask();
        auto imm_lo = eval.next_trace_mask();
        auto imm_hi = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto op0_val_lo = eval.next_trace_mask();
        auto op0_val_hi = eval.next_trace_mask();
        auto op0_prev_clock_lo = eval.next_trace_mask();
        auto op0_prev_clock_hi = eval.next_trace_mask();
        auto dst_prev_val = eval.next_trace_mask();
        auto dst_prev_clock = eval.next_trace_mask();
        auto borrow_lo = eval.next_trace_mask();
        auto borrow_hi = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(borrow_lo * (one - borrow_lo));
        eval.add_constraint(borrow_hi * (one - borrow_hi));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src_off, imm_lo, imm_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src_off, imm_lo, imm_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc + one, inst_prev_clock, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc + one, clock, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src_off, op0_prev_clock_lo, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src_off, clock, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src_off + one, op0_prev_clock_hi, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src_off + one, clock, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock, dst_prev_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, one - borrow_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {imm_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {imm_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {imm_lo - enabler + borrow_lo * two_pow_16 - op0_val_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {imm_hi - borrow_lo + borrow_hi * two_pow_16 - op0_val_hi});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto enabler = t.enabler();
        auto one = t.f_const(1);
        auto imm_lo = t.in(IN_INST0 + 2), imm_hi = t.in(IN_INST0 + 3);
        auto op0_lo = t.in(in_acc(0, ACC_VALUE)), op0_hi = t.in(in_acc(1, ACC_VALUE));
        // borrows of imm - op0 - enabler (u32_store_lt_fp_imm.rs:215-252): x < y + z
        auto borrow_lo = one - t.f_le(op0_lo + enabler, imm_lo);
        auto borrow_hi = one - t.f_le(op0_hi + borrow_lo, imm_hi);
        t.out(0, enabler);
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, imm_lo);
        t.out(7, imm_hi);
        t.out(8, t.in(IN_INST0 + 4));
        t.out(9, op0_lo);
        t.out(10, op0_hi);
        t.out(11, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(12, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(13, t.in(in_acc(2, ACC_PREV_VALUE)));
        t.out(14, t.in(in_acc(2, ACC_PREV_CLOCK)));
        t.out(15, borrow_lo);
        t.out(16, borrow_hi);
    }
};

// ------------------------------------------------------------------ u32_store_eq_fp_fp
//   .../opcodes/u32_store_eq_fp_fp.rs:200-353, :498-742
// As in the reference, dst_off is taken from instruction word 4 (u32_store_eq_fp_fp.rs:210), which a one-word
// instruction leaves at 0: only `U32StoreEqFpFp` instructions with dst_off = 0 are provable there and here.
struct U32StoreEqFpFpEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 22;
    static const char* name() { return "u32_store_eq_fp_fp"; }
    static std::vector<u32> opcodes() { return {OP_U32_STORE_EQ_FP_FP}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto opcode_constant = eval.f_const(OP_U32_STORE_EQ_FP_FP);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto src0_off = eval.next_trace_mask();
        auto src1_off = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto op0_val_lo = eval.next_trace_mask();
        auto op0_val_hi = eval.next_trace_mask();
        auto op0_prev_clock_lo = eval.next_trace_mask();
        auto op0_prev_clock_hi = eval.next_trace_mask();
        auto op1_val_lo = eval.next_trace_mask();
        auto op1_val_hi = eval.next_trace_mask();
        auto op1_prev_clock_lo = eval.next_trace_mask();
        auto op1_prev_clock_hi = eval.next_trace_mask();
        auto dst_prev_val = eval.next_trace_mask();
        auto dst_prev_clock = eval.next_trace_mask();
        auto diff_inv_lo = eval.next_trace_mask();
        auto diff_inv_hi = eval.next_trace_mask();
        auto is_eq_lo = eval.next_trace_mask();
        auto is_eq_prod = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        auto diff_lo = op1_val_lo - op0_val_lo;
        auto diff_hi = op1_val_hi - op0_val_hi;
        eval.add_constraint(diff_lo * (diff_inv_lo * diff_lo - one));
        eval.add_constraint(diff_hi * (diff_inv_hi * diff_hi - one));
        eval.add_constraint(is_eq_lo - (one - diff_lo * diff_inv_lo));
        eval.add_constraint(is_eq_prod - is_eq_lo * (one - diff_hi * diff_inv_hi));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src0_off, src1_off, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src0_off, src1_off, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off, op0_prev_clock_lo, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off, clock, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off + one, op0_prev_clock_hi, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off + one, clock, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off, op1_prev_clock_lo, op1_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off, clock, op1_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off + one, op1_prev_clock_hi, op1_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off + one, clock, op1_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock, dst_prev_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, is_eq_prod});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op1_val_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op1_val_hi});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op1_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op1_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto one = t.f_const(1);
        auto op0_lo = t.in(in_acc(0, ACC_VALUE)), op0_hi = t.in(in_acc(1, ACC_VALUE));
        auto op1_lo = t.in(in_acc(2, ACC_VALUE)), op1_hi = t.in(in_acc(3, ACC_VALUE));
        auto diff_lo = op1_lo - op0_lo, diff_hi = op1_hi - op0_hi;
        auto diff_inv_lo = t.f_inv(diff_lo), diff_inv_hi = t.f_inv(diff_hi);  // 0 -> 0
        auto is_eq_lo = one - diff_lo * diff_inv_lo;
        auto is_eq_prod = is_eq_lo * (one - diff_hi * diff_inv_hi);
        t.out(0, t.enabler());
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, t.in(IN_INST0 + 2));
        t.out(7, t.in(IN_INST0 + 4));  // sic: instruction word 4 (see the note above the struct)
        t.out(8, op0_lo);
        t.out(9, op0_hi);
        t.out(10, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(11, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(12, op1_lo);
        t.out(13, op1_hi);
        t.out(14, t.in(in_acc(2, ACC_PREV_CLOCK)));
        t.out(15, t.in(in_acc(3, ACC_PREV_CLOCK)));
        t.out(16, t.in(in_acc(4, ACC_PREV_VALUE)));
        t.out(17, t.in(in_acc(4, ACC_PREV_CLOCK)));
        t.out(18, diff_inv_lo);
        t.out(19, diff_inv_hi);
        t.out(20, is_eq_lo);
        t.out(21, is_eq_prod);
    }
};

// ------------------------------------------------------------------ u32_store_eq_fp_imm
//   .../opcodes/u32_store_eq_fp_imm.rs:191-294, :438-628
// Restated as written: the second instruction word is looked up at `pc` (not pc + 1) with the value dst_off
// (u32_store_eq_fp_imm.rs:500-512), so the Memory relation cannot balance for an enabled row; in the reference as
// here the component only ever proves its padding (it is part of every proof with 16 empty rows).  The host VM does
// not execute this opcode.
struct U32StoreEqFpImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 16;
    static const char* name() { return "u32_store_eq_fp_imm"; }
    static std::vector<u32> opcodes() { return {OP_U32_STORE_EQ_FP_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto two_pow_16 = eval.f_const(1u << 16);
        auto opcode_constant = eval.f_const(OP_U32_STORE_EQ_FP_IMM);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto src0_off = eval.next_trace_mask();
        auto imm_lo = eval.next_trace_mask();
        auto imm_hi = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto op0_val_lo = eval.next_trace_mask();
        auto op0_val_hi = eval.next_trace_mask();
        auto op0_prev_clock_lo = eval.next_trace_mask();
        auto op0_prev_clock_hi = eval.next_trace_mask();
        auto dst_prev_val = eval.next_trace_mask();
        auto dst_prev_clock = eval.next_trace_mask();
        auto diff_inv = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        auto diff = op0_val_lo + op0_val_hi * two_pow_16 - imm_lo - imm_hi * two_pow_16;
        eval.add_constraint(diff * (diff_inv * diff - one));
        eval.add_constraint(diff_inv * (diff_inv * diff - one));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src0_off, imm_lo, imm_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src0_off, imm_lo, imm_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off, op0_prev_clock_lo, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off, clock, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off + one, op0_prev_clock_hi, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off + one, clock, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock, dst_prev_val});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, one - diff * diff_inv});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {op0_val_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {imm_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {imm_hi});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto two_pow_16 = t.f_const(1u << 16);
        auto imm_lo = t.in(IN_INST0 + 2), imm_hi = t.in(IN_INST0 + 3);
        auto op0_lo = t.in(in_acc(0, ACC_VALUE)), op0_hi = t.in(in_acc(1, ACC_VALUE));
        auto diff = op0_lo + op0_hi * two_pow_16 - imm_lo - imm_hi * two_pow_16;
        t.out(0, t.enabler());
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, t.in(IN_INST0 + 1));
        t.out(6, imm_lo);
        t.out(7, imm_hi);
        t.out(8, t.in(IN_INST0 + 4));
        t.out(9, op0_lo);
        t.out(10, op0_hi);
        t.out(11, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(12, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(13, t.in(in_acc(2, ACC_PREV_VALUE)));
        t.out(14, t.in(in_acc(2, ACC_PREV_CLOCK)));
        t.out(15, t.f_inv(diff));
    }
};

// ------------------------------------------------------------------ u32_store_mul_fp_fp / u32_store_mul_fp_imm
//   .../opcodes/u32_store_mul_fp_fp.rs:235-495, :672-1058;  .../opcodes/u32_store_mul_fp_imm.rs:229-451, :651-985
// Schoolbook product of 8-bit limbs, low 32 bits kept; carries bounded through RangeCheck16(MAX_CARRY_k - carry_k).
// As in the reference, op1's previous clocks are not range-checked in the fp_fp form (5 RangeCheck20 lookups).
constexpr u32 U32_MUL_MAX_CARRY[4] = {254, 509, 764, 1019};
template <bool IMM>
struct U32StoreMulEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = IMM ? 29 : 32;
    static const char* name() { return IMM ? "u32_store_mul_fp_imm" : "u32_store_mul_fp_fp"; }
    static std::vector<u32> opcodes() { return {IMM ? OP_U32_STORE_MUL_FP_IMM : OP_U32_STORE_MUL_FP_FP}; }
    template <class E>
    void evaluate(E& eval) const {
        typedef decltype(eval.f_const(0)) F;
        auto one = eval.f_const(1);
        auto two_pow_8 = eval.f_const(1u << 8);
        auto opcode_constant = eval.f_const(IMM ? OP_U32_STORE_MUL_FP_IMM : OP_U32_STORE_MUL_FP_FP);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto src0_off = eval.next_trace_mask();
        F x[4] = {one, one, one, one}, y[4] = {one, one, one, one}, res[4] = {one, one, one, one}, carry[4] = {one, one, one, one};
        F src1_off = one, dst_off = one, op1_prev_clock_lo = one, op1_prev_clock_hi = one;
        if (IMM) {
            for (int k = 0; k < 4; k++) y[k] = eval.next_trace_mask();  // imm_0..3
            dst_off = eval.next_trace_mask();
        } else {
            src1_off = eval.next_trace_mask();
            dst_off = eval.next_trace_mask();
        }
        for (int k = 0; k < 4; k++) x[k] = eval.next_trace_mask();  // op0_0..3
        auto op0_prev_clock_lo = eval.next_trace_mask();
        auto op0_prev_clock_hi = eval.next_trace_mask();
        if (!IMM) {
            for (int k = 0; k < 4; k++) y[k] = eval.next_trace_mask();  // op1_0..3
            op1_prev_clock_lo = eval.next_trace_mask();
            op1_prev_clock_hi = eval.next_trace_mask();
        }
        auto dst_prev_val_lo = eval.next_trace_mask();
        auto dst_prev_val_hi = eval.next_trace_mask();
        auto dst_prev_clock_lo = eval.next_trace_mask();
        auto dst_prev_clock_hi = eval.next_trace_mask();
        for (int k = 0; k < 4; k++) res[k] = eval.next_trace_mask();
        for (int k = 0; k < 4; k++) carry[k] = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(enabler * (res[0] - (x[0] * y[0] - carry[0] * two_pow_8)));
        eval.add_constraint(enabler * (res[1] - (x[0] * y[1] + x[1] * y[0] + carry[0] - carry[1] * two_pow_8)));
        eval.add_constraint(enabler * (res[2] - (x[0] * y[2] + x[1] * y[1] + x[2] * y[0] + carry[1] - carry[2] * two_pow_8)));
        eval.add_constraint(enabler * (res[3] - (x[0] * y[3] + x[1] * y[2] + x[2] * y[1] + x[3] * y[0] + carry[2] - carry[3] * two_pow_8)));
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        if (IMM) {
            eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one + one, fp, clock + one});
            eval.add_to_relation(REL_MEMORY, -eval.ef(enabler),
                                 {pc, inst_prev_clock, opcode_constant, src0_off, y[0] + y[1] * two_pow_8, y[2] + y[3] * two_pow_8});
            eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src0_off, y[0] + y[1] * two_pow_8, y[2] + y[3] * two_pow_8});
            eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc + one, inst_prev_clock, dst_off});
            eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc + one, clock, dst_off});
        } else {
            eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one, fp, clock + one});
            eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src0_off, src1_off, dst_off});
            eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src0_off, src1_off, dst_off});
        }
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off, op0_prev_clock_lo, x[0] + x[1] * two_pow_8});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off, clock, x[0] + x[1] * two_pow_8});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off + one, op0_prev_clock_hi, x[2] + x[3] * two_pow_8});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off + one, clock, x[2] + x[3] * two_pow_8});
        if (!IMM) {
            eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off, op1_prev_clock_lo, y[0] + y[1] * two_pow_8});
            eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off, clock, y[0] + y[1] * two_pow_8});
            eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off + one, op1_prev_clock_hi, y[2] + y[3] * two_pow_8});
            eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off + one, clock, y[2] + y[3] * two_pow_8});
        }
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock_lo, dst_prev_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, res[0] + res[1] * two_pow_8});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off + one, dst_prev_clock_hi, dst_prev_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off + one, clock, res[2] + res[3] * two_pow_8});
        for (int k = 0; k < 4; k++) eval.add_to_relation(REL_RC8, -eval.ef_one(), {x[k]});
        for (int k = 0; k < 4; k++) eval.add_to_relation(REL_RC8, -eval.ef_one(), {y[k]});
        for (int k = 0; k < 4; k++) eval.add_to_relation(REL_RC8, -eval.ef_one(), {res[k]});
        for (int k = 0; k < 4; k++) eval.add_to_relation(REL_RC16, -eval.ef_one(), {eval.f_const(U32_MUL_MAX_CARRY[k]) - carry[k]});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_hi - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        typedef decltype(t.f_const(0)) F;
        auto two_pow_8 = t.f_const(1u << 8);
        auto lo8 = [&](F v) { return t.f_and(v, 0xff); };
        auto hi8 = [&](F v) { return t.f_shr(v, 8); };  // no mask, as in decompose_8 of the mul components
        const int dst_acc = IMM ? 2 : 4;
        F op0_lo = t.in(in_acc(0, ACC_VALUE)), op0_hi = t.in(in_acc(1, ACC_VALUE));
        F op1_lo = IMM ? t.in(IN_INST0 + 2) : t.in(in_acc(2, ACC_VALUE));
        F op1_hi = IMM ? t.in(IN_INST0 + 3) : t.in(in_acc(3, ACC_VALUE));
        F x[4] = {lo8(op0_lo), hi8(op0_lo), lo8(op0_hi), hi8(op0_hi)};
        F y[4] = {lo8(op1_lo), hi8(op1_lo), lo8(op1_hi), hi8(op1_hi)};
        F sum0 = x[0] * y[0];
        F carry0 = t.f_shr(sum0, 8);
        F sum1 = x[0] * y[1] + x[1] * y[0] + carry0;
        F carry1 = t.f_shr(sum1, 8);
        F sum2 = x[0] * y[2] + x[1] * y[1] + x[2] * y[0] + carry1;
        F carry2 = t.f_shr(sum2, 8);
        F sum3 = x[0] * y[3] + x[1] * y[2] + x[2] * y[1] + x[3] * y[0] + carry2;
        F carry3 = t.f_shr(sum3, 8);
        F res[4] = {sum0 - carry0 * two_pow_8, sum1 - carry1 * two_pow_8, sum2 - carry2 * two_pow_8, sum3 - carry3 * two_pow_8};
        F carry[4] = {carry0, carry1, carry2, carry3};
        int c = 0;
        t.out(c++, t.enabler());
        t.out(c++, t.in(IN_PC));
        t.out(c++, t.in(IN_FP));
        t.out(c++, t.in(IN_CLOCK));
        t.out(c++, t.in(IN_INST_PREV_CLOCK));
        t.out(c++, t.in(IN_INST0 + 1));
        if (IMM) {
            for (int k = 0; k < 4; k++) t.out(c++, y[k]);
            t.out(c++, t.in(IN_INST0 + 4));
        } else {
            t.out(c++, t.in(IN_INST0 + 2));
            t.out(c++, t.in(IN_INST0 + 3));
        }
        for (int k = 0; k < 4; k++) t.out(c++, x[k]);
        t.out(c++, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(c++, t.in(in_acc(1, ACC_PREV_CLOCK)));
        if (!IMM) {
            for (int k = 0; k < 4; k++) t.out(c++, y[k]);
            t.out(c++, t.in(in_acc(2, ACC_PREV_CLOCK)));
            t.out(c++, t.in(in_acc(3, ACC_PREV_CLOCK)));
        }
        t.out(c++, t.in(in_acc(dst_acc, ACC_PREV_VALUE)));
        t.out(c++, t.in(in_acc(dst_acc + 1, ACC_PREV_VALUE)));
        t.out(c++, t.in(in_acc(dst_acc, ACC_PREV_CLOCK)));
        t.out(c++, t.in(in_acc(dst_acc + 1, ACC_PREV_CLOCK)));
        for (int k = 0; k < 4; k++) t.out(c++, res[k]);
        for (int k = 0; k < 4; k++) t.out(c++, carry[k]);
    }
};
typedef U32StoreMulEval<false> U32StoreMulFpFpEval;
typedef U32StoreMulEval<true> U32StoreMulFpImmEval;

// ------------------------------------------------------------------ u32_store_div_fp_fp / u32_store_div_fp_imm (DivRem)
//   .../opcodes/u32_store_div_fp_fp.rs:298-752, :977-1509;  .../opcodes/u32_store_div_fp_imm.rs:288-710, :937-1436
// n = q * d + r with r < d: q * d as a 64-bit schoolbook product of 8-bit limbs (prod_0..7), prod + r = n on 16-bit
// limbs with the upper half forced to zero, and d - r - 1 without a final borrow.
constexpr u32 U32_DIV_MAX_CARRY[7] = {254, 509, 764, 1019, 765, 510, 255};
template <bool IMM>
struct U32StoreDivEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = IMM ? 51 : 54;
    static const char* name() { return IMM ? "u32_store_div_fp_imm" : "u32_store_div_fp_fp"; }
    static std::vector<u32> opcodes() { return {IMM ? OP_U32_STORE_DIV_REM_FP_IMM : OP_U32_STORE_DIV_REM_FP_FP}; }
    template <class E>
    void evaluate(E& eval) const {
        typedef decltype(eval.f_const(0)) F;
        auto one = eval.f_const(1);
        auto two_pow_8 = eval.f_const(1u << 8);
        auto two_pow_16 = eval.f_const(1u << 16);
        auto opcode_constant = eval.f_const(IMM ? OP_U32_STORE_DIV_REM_FP_IMM : OP_U32_STORE_DIV_REM_FP_FP);
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto src0_off = eval.next_trace_mask();
        F d[4] = {one, one, one, one}, q[4] = {one, one, one, one};
        F src1_off = one, dst_off = one, dst_rem_off = one, op1_prev_clock_lo = one, op1_prev_clock_hi = one;
        if (IMM) {
            for (int k = 0; k < 4; k++) d[k] = eval.next_trace_mask();  // imm_0..3
            dst_off = eval.next_trace_mask();
            dst_rem_off = eval.next_trace_mask();
        } else {
            src1_off = eval.next_trace_mask();
            dst_off = eval.next_trace_mask();
            dst_rem_off = eval.next_trace_mask();
        }
        auto n_lo = eval.next_trace_mask();
        auto n_hi = eval.next_trace_mask();
        auto op0_prev_clock_lo = eval.next_trace_mask();
        auto op0_prev_clock_hi = eval.next_trace_mask();
        if (!IMM) {
            for (int k = 0; k < 4; k++) d[k] = eval.next_trace_mask();  // op1_val_0..3
            op1_prev_clock_lo = eval.next_trace_mask();
            op1_prev_clock_hi = eval.next_trace_mask();
        }
        auto dst_prev_val_lo = eval.next_trace_mask();
        auto dst_prev_val_hi = eval.next_trace_mask();
        auto dst_prev_clock_lo = eval.next_trace_mask();
        auto dst_prev_clock_hi = eval.next_trace_mask();
        auto dst_rem_prev_val_lo = eval.next_trace_mask();
        auto dst_rem_prev_val_hi = eval.next_trace_mask();
        auto dst_rem_prev_clock_lo = eval.next_trace_mask();
        auto dst_rem_prev_clock_hi = eval.next_trace_mask();
        for (int k = 0; k < 4; k++) q[k] = eval.next_trace_mask();
        std::vector<F> mc, prod;
        for (int k = 0; k < 7; k++) mc.push_back(eval.next_trace_mask());    // mul_carry_0..6
        for (int k = 0; k < 8; k++) prod.push_back(eval.next_trace_mask());  // prod_0..7
        auto add_carry_0 = eval.next_trace_mask();
        auto add_carry_1 = eval.next_trace_mask();
        auto add_carry_2 = eval.next_trace_mask();
        auto add_carry_3 = eval.next_trace_mask();
        auto sub_borrow_0 = eval.next_trace_mask();
        auto sub_borrow_1 = eval.next_trace_mask();
        auto r_lo = eval.next_trace_mask();
        auto r_hi = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(enabler * add_carry_0 * (one - add_carry_0));
        eval.add_constraint(enabler * add_carry_1 * (one - add_carry_1));
        eval.add_constraint(enabler * add_carry_2 * (one - add_carry_2));
        eval.add_constraint(enabler * sub_borrow_0 * (one - sub_borrow_0));
        auto d_lo = d[0] + d[1] * two_pow_8;
        auto d_hi = d[2] + d[3] * two_pow_8;
        eval.add_constraint(enabler * (q[0] * d[0] - mc[0] * two_pow_8 - prod[0]));
        eval.add_constraint(enabler * (q[0] * d[1] + q[1] * d[0] + mc[0] - mc[1] * two_pow_8 - prod[1]));
        eval.add_constraint(enabler * (q[0] * d[2] + q[2] * d[0] + q[1] * d[1] + mc[1] - mc[2] * two_pow_8 - prod[2]));
        eval.add_constraint(enabler * (q[0] * d[3] + q[3] * d[0] + q[1] * d[2] + q[2] * d[1] + mc[2] - mc[3] * two_pow_8 - prod[3]));
        eval.add_constraint(enabler * (q[1] * d[3] + q[3] * d[1] + q[2] * d[2] + mc[3] - mc[4] * two_pow_8 - prod[4]));
        eval.add_constraint(enabler * (q[2] * d[3] + q[3] * d[2] + mc[4] - mc[5] * two_pow_8 - prod[5]));
        eval.add_constraint(enabler * (q[3] * d[3] + mc[5] - mc[6] * two_pow_8 - prod[6]));
        eval.add_constraint(enabler * (mc[6] - prod[7]));
        eval.add_constraint(enabler * (n_lo - (prod[0] + prod[1] * two_pow_8 + r_lo - add_carry_0 * two_pow_16)));
        eval.add_constraint(enabler * (n_hi - (prod[2] + prod[3] * two_pow_8 + r_hi + add_carry_0 - add_carry_1 * two_pow_16)));
        eval.add_constraint(enabler * (prod[4] + prod[5] * two_pow_8 + add_carry_1 - add_carry_2 * two_pow_16));
        eval.add_constraint(enabler * (prod[6] + prod[7] * two_pow_8 + add_carry_2 - add_carry_3 * two_pow_16));
        eval.add_constraint(enabler * add_carry_3);
        auto sub_check_lo = d[0] + d[1] * two_pow_8 + sub_borrow_0 * two_pow_16 - r_lo - one;
        auto sub_check_hi = d[2] + d[3] * two_pow_8 + sub_borrow_1 * two_pow_16 - r_hi - sub_borrow_0;
        eval.add_constraint(enabler * sub_borrow_1);
        auto res_lo = q[0] + q[1] * two_pow_8;
        auto res_hi = q[2] + q[3] * two_pow_8;
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one + one, fp, clock + one});
        if (IMM) {
            eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src0_off, d_lo, d_hi});
            eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src0_off, d_lo, d_hi});
            eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc + one, inst_prev_clock, dst_off, dst_rem_off});
            eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc + one, clock, dst_off, dst_rem_off});
        } else {
            eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src0_off, src1_off, dst_off});
            eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src0_off, src1_off, dst_off});
            eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc + one, inst_prev_clock, dst_rem_off});
            eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc + one, clock, dst_rem_off});
        }
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off, op0_prev_clock_lo, n_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off, clock, n_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src0_off + one, op0_prev_clock_hi, n_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src0_off + one, clock, n_hi});
        if (!IMM) {
            eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off, op1_prev_clock_lo, d_lo});
            eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off, clock, d_lo});
            eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src1_off + one, op1_prev_clock_hi, d_hi});
            eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src1_off + one, clock, d_hi});
        }
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock_lo, dst_prev_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, res_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off + one, dst_prev_clock_hi, dst_prev_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off + one, clock, res_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_rem_off, dst_rem_prev_clock_lo, dst_rem_prev_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_rem_off, clock, r_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_rem_off + one, dst_rem_prev_clock_hi, dst_rem_prev_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_rem_off + one, clock, r_hi});
        for (int k = 0; k < 4; k++) eval.add_to_relation(REL_RC8, -eval.ef_one(), {d[k]});
        for (int k = 0; k < 4; k++) eval.add_to_relation(REL_RC8, -eval.ef_one(), {q[k]});
        for (int k = 0; k < 8; k++) eval.add_to_relation(REL_RC8, -eval.ef_one(), {prod[k]});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {n_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {n_hi});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {r_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {r_hi});
        for (int k = 0; k < 7; k++) eval.add_to_relation(REL_RC16, -eval.ef_one(), {eval.f_const(U32_DIV_MAX_CARRY[k]) - mc[k]});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {sub_check_lo});
        eval.add_to_relation(REL_RC16, -eval.ef_one(), {sub_check_hi});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_rem_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_rem_prev_clock_hi - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        typedef decltype(t.f_const(0)) F;
        auto one = t.f_const(1);
        auto two_pow_8 = t.f_const(1u << 8);
        auto limb_max = t.f_const(0xffff);
        auto lo8 = [&](F v) { return t.f_and(v, 0xff); };
        auto hi8 = [&](F v) { return t.f_and(t.f_shr(v, 8), 0xff); };
        const int dst_acc = IMM ? 2 : 4;
        F n_lo = t.in(in_acc(0, ACC_VALUE)), n_hi = t.in(in_acc(1, ACC_VALUE));
        F d_lo = IMM ? t.in(IN_INST0 + 2) : t.in(in_acc(2, ACC_VALUE));
        F d_hi = IMM ? t.in(IN_INST0 + 3) : t.in(in_acc(3, ACC_VALUE));
        F d[4] = {lo8(d_lo), hi8(d_lo), lo8(d_hi), hi8(d_hi)};
        F q_lo = t.f_u32_divrem(n_lo, n_hi, d_lo, d_hi, 0), q_hi = t.f_u32_divrem(n_lo, n_hi, d_lo, d_hi, 1);
        F r_lo = t.f_u32_divrem(n_lo, n_hi, d_lo, d_hi, 2), r_hi = t.f_u32_divrem(n_lo, n_hi, d_lo, d_hi, 3);
        F q[4] = {lo8(q_lo), hi8(q_lo), lo8(q_hi), hi8(q_hi)};
        F raw[7] = {q[0] * d[0],
                    q[0] * d[1] + q[1] * d[0],
                    q[0] * d[2] + q[2] * d[0] + q[1] * d[1],
                    q[0] * d[3] + q[3] * d[0] + q[1] * d[2] + q[2] * d[1],
                    q[1] * d[3] + q[3] * d[1] + q[2] * d[2],
                    q[2] * d[3] + q[3] * d[2],
                    q[3] * d[3]};
        std::vector<F> mc, prod;
        for (int k = 0; k < 7; k++) {
            F with_carry = k == 0 ? raw[0] : raw[k] + mc[k - 1];
            mc.push_back(t.f_shr(with_carry, 8));
            prod.push_back(with_carry - mc[k] * two_pow_8);
        }
        prod.push_back(mc[6]);
        F add_carry_0 = one - t.f_le(prod[0] + prod[1] * two_pow_8 + r_lo, limb_max);
        F add_carry_1 = one - t.f_le(prod[2] + prod[3] * two_pow_8 + r_hi + add_carry_0, limb_max);
        F add_carry_2 = one - t.f_le(prod[4] + prod[5] * two_pow_8 + add_carry_1, limb_max);
        F add_carry_3 = one - t.f_le(prod[6] + prod[7] * two_pow_8 + add_carry_2, limb_max);
        // borrows of d - r - 1: d_val < r + 1, then d_val < r + borrow
        F sub_borrow_0 = one - t.f_le(r_lo + one, d[0] + d[1] * two_pow_8);
        F sub_borrow_1 = one - t.f_le(r_hi + sub_borrow_0, d[2] + d[3] * two_pow_8);
        int c = 0;
        t.out(c++, t.enabler());
        t.out(c++, t.in(IN_PC));
        t.out(c++, t.in(IN_FP));
        t.out(c++, t.in(IN_CLOCK));
        t.out(c++, t.in(IN_INST_PREV_CLOCK));
        t.out(c++, t.in(IN_INST0 + 1));
        if (IMM) {
            for (int k = 0; k < 4; k++) t.out(c++, d[k]);
            t.out(c++, t.in(IN_INST0 + 4));
            // dst_rem_off is rebuilt from the remainder's write address (u32_store_div_fp_imm.rs:311-313)
            t.out(c++, t.in(in_acc(4, ACC_ADDRESS)) - t.in(IN_FP));
        } else {
            t.out(c++, t.in(IN_INST0 + 2));
            t.out(c++, t.in(IN_INST0 + 3));
            t.out(c++, t.in(IN_INST0 + 4));
        }
        t.out(c++, n_lo);
        t.out(c++, n_hi);
        t.out(c++, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(c++, t.in(in_acc(1, ACC_PREV_CLOCK)));
        if (!IMM) {
            for (int k = 0; k < 4; k++) t.out(c++, d[k]);
            t.out(c++, t.in(in_acc(2, ACC_PREV_CLOCK)));
            t.out(c++, t.in(in_acc(3, ACC_PREV_CLOCK)));
        }
        t.out(c++, t.in(in_acc(dst_acc, ACC_PREV_VALUE)));
        t.out(c++, t.in(in_acc(dst_acc + 1, ACC_PREV_VALUE)));
        t.out(c++, t.in(in_acc(dst_acc, ACC_PREV_CLOCK)));
        t.out(c++, t.in(in_acc(dst_acc + 1, ACC_PREV_CLOCK)));
        t.out(c++, t.in(in_acc(dst_acc + 2, ACC_PREV_VALUE)));
        t.out(c++, t.in(in_acc(dst_acc + 3, ACC_PREV_VALUE)));
        t.out(c++, t.in(in_acc(dst_acc + 2, ACC_PREV_CLOCK)));
        t.out(c++, t.in(in_acc(dst_acc + 3, ACC_PREV_CLOCK)));
        for (int k = 0; k < 4; k++) t.out(c++, q[k]);
        for (int k = 0; k < 7; k++) t.out(c++, mc[k]);
        for (int k = 0; k < 8; k++) t.out(c++, prod[k]);
        t.out(c++, add_carry_0);
        t.out(c++, add_carry_1);
        t.out(c++, add_carry_2);
        t.out(c++, add_carry_3);
        t.out(c++, sub_borrow_0);
        t.out(c++, sub_borrow_1);
        t.out(c++, r_lo);
        t.out(c++, r_hi);
    }
};
typedef U32StoreDivEval<false> U32StoreDivFpFpEval;
typedef U32StoreDivEval<true> U32StoreDivFpImmEval;

// ------------------------------------------------------------------ u32_store_bitwise_fp_imm (And / Or / Xor with an immediate)
//   .../opcodes/u32_store_bitwise_fp_imm.rs:185-341, :500-716
struct U32StoreBitwiseFpImmEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 26;
    static const char* name() { return "u32_store_bitwise_fp_imm"; }
    static std::vector<u32> opcodes() { return {OP_U32_STORE_AND_FP_IMM, OP_U32_STORE_OR_FP_IMM, OP_U32_STORE_XOR_FP_IMM}; }
    template <class E>
    void evaluate(E& eval) const {
        auto enabler = eval.next_trace_mask();
        auto pc = eval.next_trace_mask();
        auto fp = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto inst_prev_clock = eval.next_trace_mask();
        auto opcode_constant = eval.next_trace_mask();
        auto src_off = eval.next_trace_mask();
        auto imm_0 = eval.next_trace_mask();
        auto imm_1 = eval.next_trace_mask();
        auto imm_2 = eval.next_trace_mask();
        auto imm_3 = eval.next_trace_mask();
        auto dst_off = eval.next_trace_mask();
        auto op0_val_0 = eval.next_trace_mask();
        auto op0_val_1 = eval.next_trace_mask();
        auto op0_val_2 = eval.next_trace_mask();
        auto op0_val_3 = eval.next_trace_mask();
        auto op0_prev_clock_lo = eval.next_trace_mask();
        auto op0_prev_clock_hi = eval.next_trace_mask();
        auto dst_prev_val_lo = eval.next_trace_mask();
        auto dst_prev_val_hi = eval.next_trace_mask();
        auto dst_val_0 = eval.next_trace_mask();
        auto dst_val_1 = eval.next_trace_mask();
        auto dst_val_2 = eval.next_trace_mask();
        auto dst_val_3 = eval.next_trace_mask();
        auto dst_prev_clock_lo = eval.next_trace_mask();
        auto dst_prev_clock_hi = eval.next_trace_mask();
        auto two_pow_8 = eval.f_const(1u << 8);
        auto one = eval.f_const(1);
        eval.add_constraint(enabler * (enabler - one));
        auto bitwise_op = opcode_constant - eval.f_const(OP_U32_STORE_AND_FP_IMM);
        auto op0_val_lo = op0_val_0 + op0_val_1 * two_pow_8;
        auto op0_val_hi = op0_val_2 + op0_val_3 * two_pow_8;
        auto imm_lo = imm_0 + imm_1 * two_pow_8;
        auto imm_hi = imm_2 + imm_3 * two_pow_8;
        auto dst_val_lo = dst_val_0 + dst_val_1 * two_pow_8;
        auto dst_val_hi = dst_val_2 + dst_val_3 * two_pow_8;
        eval.add_to_relation(REL_REGISTERS, -eval.ef(enabler), {pc, fp, clock});
        eval.add_to_relation(REL_REGISTERS, eval.ef(enabler), {pc + one + one, fp, clock + one});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc, inst_prev_clock, opcode_constant, src_off, imm_lo, imm_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc, clock, opcode_constant, src_off, imm_lo, imm_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {pc + one, inst_prev_clock, dst_off});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {pc + one, clock, dst_off});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src_off, op0_prev_clock_lo, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src_off, clock, op0_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + src_off + one, op0_prev_clock_hi, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + src_off + one, clock, op0_val_hi});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off, dst_prev_clock_lo, dst_prev_val_lo});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off, clock, dst_val_lo});
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {fp + dst_off + one, dst_prev_clock_hi, dst_prev_val_hi});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {fp + dst_off + one, clock, dst_val_hi});
        eval.add_to_relation(REL_BITWISE, -eval.ef_one(), {bitwise_op, op0_val_0, imm_0, dst_val_0});
        eval.add_to_relation(REL_BITWISE, -eval.ef_one(), {bitwise_op, op0_val_1, imm_1, dst_val_1});
        eval.add_to_relation(REL_BITWISE, -eval.ef_one(), {bitwise_op, op0_val_2, imm_2, dst_val_2});
        eval.add_to_relation(REL_BITWISE, -eval.ef_one(), {bitwise_op, op0_val_3, imm_3, dst_val_3});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - inst_prev_clock - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - op0_prev_clock_hi - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_lo - enabler});
        eval.add_to_relation(REL_RC20, -eval.ef_one(), {clock - dst_prev_clock_hi - enabler});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        auto enabler = t.enabler();
        auto base_opcode = t.f_const(OP_U32_STORE_AND_FP_IMM);
        // padding rows (default bundle = Ret) are rewritten to U32_STORE_AND_FP_IMM (u32_store_bitwise_fp_imm.rs:192-198)
        auto opcode_constant = enabler * (t.in(IN_INST0) - base_opcode) + base_opcode;
        auto lo8 = [&](decltype(enabler) v) { return t.f_and(v, 0xff); };
        auto hi8 = [&](decltype(enabler) v) { return t.f_and(t.f_shr(v, 8), 0xff); };
        auto imm_lo = t.in(IN_INST0 + 2), imm_hi = t.in(IN_INST0 + 3);
        auto op0_lo = t.in(in_acc(0, ACC_VALUE)), op0_hi = t.in(in_acc(1, ACC_VALUE));
        auto dst_lo = t.in(in_acc(2, ACC_VALUE)), dst_hi = t.in(in_acc(3, ACC_VALUE));
        t.out(0, enabler);
        t.out(1, t.in(IN_PC));
        t.out(2, t.in(IN_FP));
        t.out(3, t.in(IN_CLOCK));
        t.out(4, t.in(IN_INST_PREV_CLOCK));
        t.out(5, opcode_constant);
        t.out(6, t.in(IN_INST0 + 1));
        t.out(7, lo8(imm_lo));
        t.out(8, hi8(imm_lo));
        t.out(9, lo8(imm_hi));
        t.out(10, hi8(imm_hi));
        t.out(11, t.in(IN_INST0 + 4));
        t.out(12, lo8(op0_lo));
        t.out(13, hi8(op0_lo));
        t.out(14, lo8(op0_hi));
        t.out(15, hi8(op0_hi));
        t.out(16, t.in(in_acc(0, ACC_PREV_CLOCK)));
        t.out(17, t.in(in_acc(1, ACC_PREV_CLOCK)));
        t.out(18, t.in(in_acc(2, ACC_PREV_VALUE)));
        t.out(19, t.in(in_acc(3, ACC_PREV_VALUE)));
        t.out(20, lo8(dst_lo));
        t.out(21, hi8(dst_lo));
        t.out(22, lo8(dst_hi));
        t.out(23, hi8(dst_hi));
        t.out(24, t.in(in_acc(2, ACC_PREV_CLOCK)));
        t.out(25, t.in(in_acc(3, ACC_PREV_CLOCK)));
    }
};

// Opcode components in claim order (crates/prover/src/components/opcodes/mod.rs:223-268): all 26.
#define CM31_OPCODE_EVALS(X)                                                                                          \
    X(AssertEqFpImmEval) X(CallAbsImmEval) X(JmpImmEval) X(JnzFpImmEval) X(RetEval) X(StoreImmEval) X(StoreFpFpEval) \
    X(StoreFpImmEval) X(DoubleDerefFpImmEval) X(DoubleDerefFpFpEval) X(StoreFramePointerEval) X(U32StoreImmEval)   \
    X(U32StoreAddFpImmEval) X(U32StoreMulFpImmEval) X(U32StoreDivFpImmEval) X(U32StoreEqFpFpEval)                  \
    X(U32StoreEqFpImmEval) X(U32StoreLtFpImmEval) X(U32StoreLtFpFpEval) X(U32StoreAddFpFpEval)                     \
    X(U32StoreSubFpFpEval) X(U32StoreMulFpFpEval) X(U32StoreDivFpFpEval) X(U32StoreBitwiseFpFpEval)                \
    X(U32StoreBitwiseFpImmEval) X(StoreLeFpImmEval)

// ------------------------------------------------------------------ memory (boundary values)
// inputs: address, clock, value0..3, multiplicity, root
struct MemoryEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 9;
    static const char* name() { return "memory"; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto m31_2 = eval.f_const(2);
        auto m31_3 = eval.f_const(3);
        auto m31_4 = eval.f_const(4);
        auto tree_height = eval.f_const(TREE_HEIGHT);
        auto enabler = eval.next_trace_mask();
        auto address = eval.next_trace_mask();
        auto clock = eval.next_trace_mask();
        auto value0 = eval.next_trace_mask();
        auto value1 = eval.next_trace_mask();
        auto value2 = eval.next_trace_mask();
        auto value3 = eval.next_trace_mask();
        auto multiplicity = eval.next_trace_mask();
        auto root = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_to_relation(REL_MEMORY, eval.ef(multiplicity), {address, clock, value0, value1, value2, value3});
        eval.add_to_relation(REL_MERKLE, -eval.ef(enabler), {address * m31_4, tree_height, value0, root});
        eval.add_to_relation(REL_MERKLE, -eval.ef(enabler), {address * m31_4 + one, tree_height, value1, root});
        eval.add_to_relation(REL_MERKLE, -eval.ef(enabler), {address * m31_4 + m31_2, tree_height, value2, root});
        eval.add_to_relation(REL_MERKLE, -eval.ef(enabler), {address * m31_4 + m31_3, tree_height, value3, root});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        t.out(0, t.enabler());
        for (int k = 0; k < 8; k++) t.out(1 + k, t.in(k));
    }
};

// ------------------------------------------------------------------ merkle (partial Merkle tree of the boundary memory)
//   crates/prover/src/components/merkle.rs:73-170 (write_trace), :285-377 (evaluate)
// inputs: index, depth, left_value, right_value, parent_value, left/right/parent multiplicity, root
struct MerkleEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 10;
    static const char* name() { return "merkle"; }
    template <class E>
    void evaluate(E& eval) const {
        auto one = eval.f_const(1);
        auto two = eval.f_const(2);
        auto m31_2_inv = eval.f_const(m31_inv(2));
        auto enabler = eval.next_trace_mask();
        auto index = eval.next_trace_mask();
        auto depth = eval.next_trace_mask();
        auto left_value = eval.next_trace_mask();
        auto right_value = eval.next_trace_mask();
        auto parent_value = eval.next_trace_mask();
        auto left_multiplicity = eval.next_trace_mask();
        auto right_multiplicity = eval.next_trace_mask();
        auto parent_multiplicity = eval.next_trace_mask();
        auto root = eval.next_trace_mask();
        eval.add_constraint(enabler * (one - enabler));
        eval.add_constraint(left_multiplicity * (left_multiplicity - one) * (left_multiplicity - one * two));
        eval.add_constraint(right_multiplicity * (right_multiplicity - one) * (right_multiplicity - one * two));
        eval.add_constraint(parent_multiplicity * (parent_multiplicity - one) * (parent_multiplicity - one * two));
        eval.add_to_relation(REL_MERKLE, eval.ef(left_multiplicity), {index, depth, left_value, root});
        eval.add_to_relation(REL_MERKLE, eval.ef(right_multiplicity), {index + one, depth, right_value, root});
        eval.add_to_relation(REL_MERKLE, -eval.ef(parent_multiplicity), {index * m31_2_inv, depth - one, parent_value, root});
        eval.add_to_relation(REL_POSEIDON2, eval.ef(enabler), {left_value, right_value});
        eval.add_to_relation(REL_POSEIDON2, -eval.ef(enabler), {parent_value});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        t.out(0, t.enabler());
        for (int k = 0; k < 9; k++) t.out(1 + k, t.in(k));
    }
};

// ------------------------------------------------------------------ poseidon2 (one permutation per row)
//   crates/prover/src/components/poseidon2.rs:143-290 (write_trace), :385-505 (evaluate); round structure in
//   csrc/cairo/poseidon2.hpp, constants in csrc/cairo/poseidon2_constants.hpp (KAT-pinned).  inputs: the 16 words of the initial state.
struct Poseidon2Eval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 1 + POSEIDON2_T * (1 + POSEIDON2_FULL_ROUNDS * 3) + 3 * POSEIDON2_PARTIAL_ROUNDS;  // 443
    static const char* name() { return "poseidon2"; }
    template <class E>
    void evaluate(E& eval) const {
        typedef decltype(eval.f_const(0)) F;
        const Poseidon2Constants& k = poseidon2_constants();
        auto enabler = eval.next_trace_mask();
        std::array<F, 16> state = {enabler, enabler, enabler, enabler, enabler, enabler, enabler, enabler,
                                   enabler, enabler, enabler, enabler, enabler, enabler, enabler, enabler};
        for (int i = 0; i < 16; i++) state[i] = eval.next_trace_mask();
        std::array<F, 16> initial_state = state;
        auto mulc = [&](F x, u32 c) { return x * eval.f_const(c); };
        auto constrain_to_mask = [&](F& s) {
            auto m = eval.next_trace_mask();
            eval.add_constraint(enabler * (s - m));
            s = m;
        };
        auto full_round = [&](int r) {
            for (int i = 0; i < 16; i++) state[i] = state[i] + eval.f_const(k.external[r][i]);
            std::array<F, 16> round_input = state;
            for (int i = 0; i < 16; i++) state[i] = state[i] * state[i];
            for (int i = 0; i < 16; i++) constrain_to_mask(state[i]);
            for (int i = 0; i < 16; i++) state[i] = state[i] * state[i];
            for (int i = 0; i < 16; i++) constrain_to_mask(state[i]);
            for (int i = 0; i < 16; i++) state[i] = state[i] * round_input[i];
            poseidon2_external_matrix(state);
            for (int i = 0; i < 16; i++) constrain_to_mask(state[i]);
        };
        poseidon2_external_matrix(state);
        for (int r = 0; r < POSEIDON2_FULL_ROUNDS / 2; r++) full_round(r);
        for (int r = 0; r < POSEIDON2_PARTIAL_ROUNDS; r++) {
            state[0] = state[0] + eval.f_const(k.internal[r]);
            F round_input = state[0];
            {
                auto m = eval.next_trace_mask();
                eval.add_constraint(enabler * (state[0] * state[0] - m));
                state[0] = m;
            }
            {
                auto m = eval.next_trace_mask();
                eval.add_constraint(enabler * (state[0] * state[0] - m));
                state[0] = m;
            }
            {
                auto m = eval.next_trace_mask();
                eval.add_constraint(enabler * (round_input * state[0] - m));
                state[0] = m;
            }
            poseidon2_internal_matrix(state, mulc);
        }
        for (int r = 0; r < POSEIDON2_FULL_ROUNDS / 2; r++) full_round(POSEIDON2_FULL_ROUNDS / 2 + r);
        eval.add_to_relation(REL_POSEIDON2, -eval.ef(enabler), std::vector<F>(initial_state.begin(), initial_state.end()));
        eval.add_to_relation(REL_POSEIDON2, eval.ef(enabler), {state[0]});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        typedef decltype(t.f_const(0)) F;
        const Poseidon2Constants& k = poseidon2_constants();
        int c = 0;
        F enabler = t.enabler();
        t.out(c++, enabler);
        std::array<F, 16> state = {enabler, enabler, enabler, enabler, enabler, enabler, enabler, enabler,
                                   enabler, enabler, enabler, enabler, enabler, enabler, enabler, enabler};
        for (int i = 0; i < 16; i++) {
            state[i] = t.in(i);
            t.out(c++, state[i]);
        }
        auto mulc = [&](F x, u32 cst) { return x * t.f_const(cst); };
        auto full_round = [&](int r) {
            for (int i = 0; i < 16; i++) state[i] = state[i] + t.f_const(k.external[r][i]);
            std::array<F, 16> round_input = state;
            for (int i = 0; i < 16; i++) state[i] = state[i] * state[i];
            for (int i = 0; i < 16; i++) t.out(c++, state[i]);
            for (int i = 0; i < 16; i++) state[i] = state[i] * state[i];
            for (int i = 0; i < 16; i++) t.out(c++, state[i]);
            for (int i = 0; i < 16; i++) state[i] = state[i] * round_input[i];
            poseidon2_external_matrix(state);
            for (int i = 0; i < 16; i++) t.out(c++, state[i]);
        };
        poseidon2_external_matrix(state);
        for (int r = 0; r < POSEIDON2_FULL_ROUNDS / 2; r++) full_round(r);
        for (int r = 0; r < POSEIDON2_PARTIAL_ROUNDS; r++) {
            state[0] = state[0] + t.f_const(k.internal[r]);
            F round_input = state[0];
            state[0] = state[0] * state[0];
            t.out(c++, state[0]);
            state[0] = state[0] * state[0];
            t.out(c++, state[0]);
            state[0] = round_input * state[0];
            t.out(c++, state[0]);
            poseidon2_internal_matrix(state, mulc);
        }
        for (int r = 0; r < POSEIDON2_FULL_ROUNDS / 2; r++) full_round(POSEIDON2_FULL_ROUNDS / 2 + r);
    }
};

// ------------------------------------------------------------------ clock_update
// inputs: addr, prev_clk, value0..3
struct ClockUpdateEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 7;
    static const char* name() { return "clock_update"; }
    template <class E>
    void evaluate(E& eval) const {
        auto enabler = eval.next_trace_mask();
        auto address = eval.next_trace_mask();
        auto prev_clk = eval.next_trace_mask();
        auto value0 = eval.next_trace_mask();
        auto value1 = eval.next_trace_mask();
        auto value2 = eval.next_trace_mask();
        auto value3 = eval.next_trace_mask();
        auto one = eval.f_const(1);
        eval.add_constraint(enabler * (one - enabler));
        eval.add_to_relation(REL_MEMORY, -eval.ef(enabler), {address, prev_clk, value0, value1, value2, value3});
        eval.add_to_relation(REL_MEMORY, eval.ef(enabler), {address, prev_clk + eval.f_const(RC20_LIMIT), value0, value1, value2, value3});
        eval.finalize_logup_in_pairs();
    }
    template <class T>
    void write_trace(T& t) const {
        t.out(0, t.enabler());
        for (int k = 0; k < 6; k++) t.out(1 + k, t.in(k));
    }
};

// ------------------------------------------------------------------ range_check_N
struct RangeCheckEval : OpcodeEvalBase {
    int relation;
    static constexpr int N_TRACE_COLUMNS = 1;
    static const char* name() { return "range_check"; }
    long cache_tag() const { return relation; }  // the captured graph names the relation's parameters
    std::string column_id() const { return "range_check_" + std::to_string(log_size_); }
    template <class E>
    void evaluate(E& eval) const {
        auto value = eval.get_preprocessed_column(column_id());
        auto multiplicity = eval.next_trace_mask();
        eval.add_to_relation(relation, eval.ef(multiplicity), {value});
        eval.finalize_logup();
    }
};

// ------------------------------------------------------------------ bitwise (table multiplicities)
// preprocessed columns (operation_id, input1, input2, result) stacked for AND, OR, XOR: 2^18 rows
struct BitwiseEval : OpcodeEvalBase {
    static constexpr int N_TRACE_COLUMNS = 1;
    static const char* name() { return "bitwise"; }
    static std::string column_id(int k) { return "bitwise_stacked_col_" + std::to_string(k); }
    template <class E>
    void evaluate(E& eval) const {
        auto operation_id = eval.get_preprocessed_column(column_id(0));
        auto input1 = eval.get_preprocessed_column(column_id(1));
        auto input2 = eval.get_preprocessed_column(column_id(2));
        auto result = eval.get_preprocessed_column(column_id(3));
        auto multiplicity = eval.next_trace_mask();
        eval.add_to_relation(REL_BITWISE, eval.ef(multiplicity), {operation_id, input1, input2, result});
        eval.finalize_logup();
    }
};
// value of preprocessed bitwise column `k` at table row `i` (preprocessed/bitwise.rs:253-290)
CM_HD u32 bitwise_table_value(int k, u32 i) {
    u32 op = i >> 16, in1 = (i >> 8) & 0xff, in2 = i & 0xff;
    if (op >= 3) return 0;  // rows beyond the three stacked operations are zero padding
    if (k == 0) return op;
    if (k == 1) return in1;
    if (k == 2) return in2;
    return op == 0 ? (in1 & in2) : op == 1 ? (in1 | in2) : (in1 ^ in2);
}

}  // namespace cm31
