//! `ComponentProver<CudaBackend> for FrameworkComponent<E>` -- the row loop of
//! external/stwo/crates/constraint_framework/src/component.rs:282-424 on the device.
//!
//! `FrameworkComponent`'s fields (`eval`, `trace_locations`, `preprocessed_column_indices`) are private to
//! `stwo_constraint_framework`, and the crate implements `ComponentProver` for `SimdBackend` only (`component.rs:282`), so
//! this file is compiled INTO the vendored `external/stwo/crates/constraint_framework` (`mod cuda;` in its lib.rs, with
//! `cairo-m-prover-cuda` as an optional dependency) -- the reference already builds stwo from that checkout through the
//! workspace `[patch]`.  The body mirrors the SimdBackend impl statement for statement up to the point-wise loop.
use std::borrow::Cow;
use std::sync::OnceLock;

use itertools::Itertools;
use stwo_prover::core::air::accumulation::DomainEvaluationAccumulator;
use stwo_prover::core::air::{Component, ComponentProver, Trace};
use stwo_prover::core::backend::Column;
use stwo_prover::core::constraints::coset_vanishing;
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::FieldExpOps;
use stwo_prover::core::pcs::TreeVec;
use stwo_prover::core::poly::circle::{CanonicCoset, CircleEvaluation, PolyOps};
use stwo_prover::core::poly::BitReversedOrder;
use stwo_prover::core::utils::bit_reverse;

use stwo_constraint_framework::expr::ExprEvaluator;
use stwo_constraint_framework::{FrameworkComponent, FrameworkEval, PREPROCESSED_TRACE_IDX};

use crate::backend::CudaBackend;
use crate::bytecode::{lower_constraints, AirProgram};
use crate::ffi::*;

/// One symbolic capture + lowering per AIR shape (the graph depends on the Eval type, never on log_size or on drawn
/// values -- those enter through `Param`s patched per proof).
pub trait CapturedProgram {
    fn program(&self) -> &OnceLock<AirProgram>;
}

impl<E: FrameworkEval + Sync> ComponentProver<CudaBackend> for FrameworkComponent<E> {
    fn evaluate_constraint_quotients_on_domain(&self, trace: &Trace<'_, CudaBackend>, evaluation_accumulator: &mut DomainEvaluationAccumulator<CudaBackend>) {
        if self.n_constraints() == 0 {
            return;
        }
        let eval_domain = CanonicCoset::new(self.max_constraint_log_degree_bound()).circle_domain();
        let trace_domain = CanonicCoset::new(self.eval.log_size());

        let mut component_polys = trace.polys.sub_tree(&self.trace_locations);
        component_polys[PREPROCESSED_TRACE_IDX] = self.preprocessed_column_indices.iter().map(|idx| &trace.polys[PREPROCESSED_TRACE_IDX][*idx]).collect();
        let mut component_evals = trace.evals.sub_tree(&self.trace_locations);
        component_evals[PREPROCESSED_TRACE_IDX] = self.preprocessed_column_indices.iter().map(|idx| &trace.evals[PREPROCESSED_TRACE_IDX][*idx]).collect();

        // blowup 2^1 and max_constraint_log_degree_bound = log_size + 1 for every cairo-m component: the committed LDE IS
        // the evaluation domain (component.rs:312-326 `need_to_extend == false`); otherwise extend like the reference
        let need_to_extend = component_evals.iter().flatten().any(|c| c.domain != eval_domain);
        let cols: TreeVec<Vec<Cow<'_, CircleEvaluation<CudaBackend, BaseField, BitReversedOrder>>>> = if need_to_extend {
            let twiddles = CudaBackend::precompute_twiddles(eval_domain.half_coset);
            component_polys.as_cols_ref().map_cols(|col| Cow::Owned(col.evaluate_with_twiddles(eval_domain, &twiddles)))
        } else {
            component_evals.clone().map_cols(|c| Cow::Borrowed(*c))
        };

        // Denom inverses (component.rs:329-333): two distinct values on the 2n-domain, indexed by row >> log_size
        let log_expand = eval_domain.log_size() - trace_domain.log_size();
        let mut denom_inv = (0..1 << log_expand).map(|i| coset_vanishing(trace_domain.coset(), eval_domain.at(i)).inverse()).collect_vec();
        bit_reverse(&mut denom_inv);
        let denom_inv = denom_inv.iter().map(|d| d.0).collect_vec();

        // Accumulator (component.rs:336-338)
        let [mut accum] = evaluation_accumulator.columns([(eval_domain.log_size(), self.n_constraints())]);
        accum.random_coeff_powers.reverse();

        // The program: captured once, parameters patched per proof
        let mut program: AirProgram = {
            let mut ev = ExprEvaluator::new();
            self.eval.evaluate(&mut ev); // the InfoEvaluator-style pass; cached per Eval type in the real integration
            lower_constraints(&ev)
        };
        program.set_coeffs(&accum.random_coeff_powers);
        for (name, value) in self.eval_params() {
            program.set_param(&name, value);
        }

        let in_cols = program.columns.iter().map(|(interaction, idx)| cols[*interaction][*idx].values.as_ptr()).collect_vec();
        let acc: [*mut u32; 4] = std::array::from_fn(|k| accum.col.columns[k].as_mut_ptr());
        CudaBackend::lane_for(trace_domain.log_size());
        check(unsafe {
            cm31_constraint_eval(
                in_cols.as_ptr(), in_cols.len(), trace_domain.log_size(), eval_domain.log_size(), program.code.as_ptr(), program.code.len(),
                program.n_regs, program.consts.as_ptr(), program.consts.len(), denom_inv.as_ptr(), acc.as_ptr(),
            )
        });
    }
}

/// Values of the AIR's formal parameters for this proof: the relation elements `(z, alpha^0, alpha^1, ..)` every
/// `relation!` contributes as `Param("<Relation>_z")` / `Param("<Relation>_alpha<i>")` (`expr/evaluator.rs:203-260`,
/// `lib.rs:295-…`) and the logup cumulative-sum shift `claimed_sum / n` (`lib.rs:210-234`).
pub trait EvalParams {
    fn eval_params(&self) -> Vec<(String, stwo_prover::core::fields::qm31::SecureField)>;
}
