//! `ExprEvaluator` -> AIR bytecode.
//!
//! The Rust `FrameworkEval::evaluate<E: EvalAtRow>` bodies cannot run on the device.  stwo already ships a symbolic
//! evaluator, `ExprEvaluator` (external/stwo/crates/constraint_framework/src/expr/evaluator.rs:63-260), that records
//! every constraint as an expression tree over `Col(interaction, idx, offset)`, constants and named parameters
//! (`expr/mod.rs:17-81`).  This module lowers those trees to the register bytecode `cm31_constraint_eval` /
//! `cm31_air_program` execute -- the same instruction set and encoding as `csrc/host/air_expr.hpp` (`AirOp`, `air_encode`):
//!
//!   word = op | dst << 8 | a << 24 | b << 44      (op 8 bits, dst 16 bits, a / b 20 bits)
//!
//! Base-field values occupy one register, secure-field values four consecutive registers.  libcm31 keys its AOT-specialised
//! kernels (csrc/generated/) by a hash of the instruction words; a program it has no generated kernel for runs on the
//! bytecode interpreter (csrc/air.cu), so any lowering that computes the right values is correct -- a lowering that
//! reproduces `ProgramBuilder::compile`'s instruction order additionally gets the fast kernels.  `cm31_air_shapes` lets the
//! shim assert that both sides agree on every component's shape before trusting either.
use std::collections::HashMap;

use stwo_constraint_framework::expr::{BaseExpr, ColumnExpr, ExprEvaluator, ExtExpr};
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;

#[repr(u8)]
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum AirOp {
    Load = 1,
    ConstF,
    ConstE,
    Add,
    Sub,
    Mul,
    Neg,
    EAdd,
    ESub,
    EMul,
    ENeg,
    EMulF,
    EAddF,
    ESubF,
    F2E,
    Mov,
    EInv,
    ConstraintE,
    ConstraintF,
    StoreE,
    StoreF,
    Hist,
    Inv,
}

pub fn encode(op: AirOp, dst: u32, a: u32, b: u32) -> u64 {
    (op as u64) | (((dst & 0xffff) as u64) << 8) | (((a & 0xfffff) as u64) << 24) | (((b & 0xfffff) as u64) << 44)
}

/// What `cm31_constraint_eval` / `cm31_air_program` take.
#[derive(Default, Clone)]
pub struct AirProgram {
    pub code: Vec<u64>,
    pub consts: Vec<u32>,
    pub n_regs: u32,
    /// `Param(name)` -> word offset of its 4 words in `consts` (filled per proof: relation z / alpha powers, claimed sum)
    pub param_slots: Vec<(String, u32)>,
    /// word offset of the k-th random-coefficient power (4 words each), filled per proof
    pub coeff_slots: Vec<u32>,
    /// (interaction, idx) of every input column, in `in_cols` order
    pub columns: Vec<(usize, usize)>,
}

#[derive(Copy, Clone)]
enum Reg {
    F(u32),
    E(u32),
}

struct Lowering {
    prog: AirProgram,
    next_reg: u32,
    col_index: HashMap<(usize, usize), u32>,
    f_cache: HashMap<String, u32>,
    params: HashMap<String, u32>,
}

impl Lowering {
    fn reg(&mut self, ext: bool) -> u32 {
        // SSA-style: one fresh register (or aligned group of four) per value.  `ProgramBuilder::compile` recycles dead
        // registers (free lists, air_expr.hpp); the interpreter supports up to 2048 registers per program.
        if ext {
            self.next_reg = (self.next_reg + 3) & !3;
        }
        let r = self.next_reg;
        self.next_reg += if ext { 4 } else { 1 };
        r
    }
    fn emit(&mut self, op: AirOp, dst: u32, a: u32, b: u32) {
        self.prog.code.push(encode(op, dst, a, b));
    }
    fn const_f(&mut self, v: BaseField) -> u32 {
        let slot = self.prog.consts.len() as u32;
        self.prog.consts.push(v.0);
        let r = self.reg(false);
        self.emit(AirOp::ConstF, r, slot, 0);
        r
    }
    fn const_e_words(&mut self, words: [u32; 4]) -> u32 {
        let slot = self.prog.consts.len() as u32;
        self.prog.consts.extend(words);
        let r = self.reg(true);
        self.emit(AirOp::ConstE, r, slot, 0);
        r
    }
    fn param(&mut self, name: &str) -> u32 {
        if let Some(r) = self.params.get(name) {
            return *r;
        }
        let slot = self.prog.consts.len() as u32;
        self.prog.consts.extend([0u32; 4]);
        self.prog.param_slots.push((name.to_string(), slot));
        let r = self.reg(true);
        self.emit(AirOp::ConstE, r, slot, 0);
        self.params.insert(name.to_string(), r);
        r
    }
    fn column(&mut self, c: &ColumnExpr) -> u32 {
        let key = (c.interaction(), c.idx());
        let n = self.col_index.len() as u32;
        let idx = *self.col_index.entry(key).or_insert_with(|| n);
        if idx == n {
            self.prog.columns.push(key);
        }
        let r = self.reg(false);
        // b = signed 20-bit mask offset in trace-domain steps (component.rs:240-251: [0] everywhere, [-1, 0] on the last
        // four logup columns); the kernel maps it with offset_bit_reversed_circle_domain_index (core/utils.rs:74-90)
        self.emit(AirOp::Load, r, idx, (c.offset() as i32 as u32) & 0xfffff);
        r
    }
    fn base(&mut self, e: &BaseExpr) -> u32 {
        let key = format!("{e:?}");
        if let Some(r) = self.f_cache.get(&key) {
            return *r; // common sub-expressions (intermediates are inlined by ExprEvaluator) are computed once
        }
        let r = match e {
            BaseExpr::Col(c) => self.column(c),
            BaseExpr::Const(v) => self.const_f(*v),
            BaseExpr::Param(name) => panic!("base-field parameter {name}: cairo-m's relations are secure-field parameters"),
            BaseExpr::Add(a, b) => self.bin(AirOp::Add, a, b),
            BaseExpr::Sub(a, b) => self.bin(AirOp::Sub, a, b),
            BaseExpr::Mul(a, b) => self.bin(AirOp::Mul, a, b),
            BaseExpr::Neg(a) => {
                let (ra, r) = (self.base(a), self.reg(false));
                self.emit(AirOp::Neg, r, ra, 0);
                r
            }
            BaseExpr::Inv(a) => {
                let (ra, r) = (self.base(a), self.reg(false));
                self.emit(AirOp::Inv, r, ra, 0);
                r
            }
        };
        self.f_cache.insert(key, r);
        r
    }
    fn bin(&mut self, op: AirOp, a: &BaseExpr, b: &BaseExpr) -> u32 {
        let (ra, rb) = (self.base(a), self.base(b));
        let r = self.reg(false);
        self.emit(op, r, ra, rb);
        r
    }
    fn ext(&mut self, e: &ExtExpr) -> Reg {
        match e {
            ExtExpr::SecureCol(coords) => {
                // from_partial_evals of four base values: assemble with MOVs into one aligned group
                let parts: Vec<u32> = coords.iter().map(|c| self.base(c)).collect();
                let r = self.reg(true);
                for (k, p) in parts.iter().enumerate() {
                    self.emit(AirOp::Mov, r + k as u32, *p, 0);
                }
                Reg::E(r)
            }
            ExtExpr::Const(v) => Reg::E(self.const_e_words(v.to_m31_array().map(|x| x.0))),
            ExtExpr::Param(name) => Reg::E(self.param(name)),
            ExtExpr::Add(a, b) => self.ext_bin(AirOp::EAdd, a, b),
            ExtExpr::Sub(a, b) => self.ext_bin(AirOp::ESub, a, b),
            ExtExpr::Mul(a, b) => self.ext_bin(AirOp::EMul, a, b),
            ExtExpr::Neg(a) => {
                let Reg::E(ra) = self.ext(a) else { unreachable!() };
                let r = self.reg(true);
                self.emit(AirOp::ENeg, r, ra, 0);
                Reg::E(r)
            }
        }
    }
    fn ext_bin(&mut self, op: AirOp, a: &ExtExpr, b: &ExtExpr) -> Reg {
        let (Reg::E(ra), Reg::E(rb)) = (self.ext(a), self.ext(b)) else { unreachable!() };
        let r = self.reg(true);
        self.emit(op, r, ra, rb);
        Reg::E(r)
    }
}

/// The constraint program of a component: `acc = Σ_k coeff_k · constraint_k` with `coeff_k` the k-th entry of the
/// component's (reversed) random-coefficient powers (`component.rs:336-338`); the kernel multiplies by the row's
/// `denom_inv` and adds into the accumulator columns (`component.rs:413-421`).
pub fn lower_constraints(ev: &ExprEvaluator) -> AirProgram {
    let mut lo = Lowering { prog: AirProgram::default(), next_reg: 0, col_index: HashMap::new(), f_cache: HashMap::new(), params: HashMap::new() };
    for constraint in &ev.constraints {
        let Reg::E(r) = lo.ext(constraint) else { unreachable!() };
        let slot = lo.prog.consts.len() as u32;
        lo.prog.consts.extend([0u32; 4]);
        lo.prog.coeff_slots.push(slot);
        lo.emit(AirOp::ConstraintE, 0, r, slot);
    }
    lo.prog.n_regs = (lo.next_reg + 3) & !3;
    lo.prog
}

impl AirProgram {
    pub fn set_param(&mut self, name: &str, v: SecureField) {
        for (n, slot) in &self.param_slots {
            if n == name {
                let w = v.to_m31_array().map(|x| x.0);
                self.consts[*slot as usize..*slot as usize + 4].copy_from_slice(&w);
            }
        }
    }
    pub fn set_coeffs(&mut self, powers: &[SecureField]) {
        assert_eq!(powers.len(), self.coeff_slots.len());
        for (slot, p) in self.coeff_slots.iter().zip(powers) {
            let w = p.to_m31_array().map(|x| x.0);
            self.consts[*slot as usize..*slot as usize + 4].copy_from_slice(&w);
        }
    }
}
