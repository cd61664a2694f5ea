//! `QuotientOps for CudaBackend` (external/stwo/crates/prover/src/core/pcs/quotients.rs:22-35); replaces
//! `simd/quotients.rs:33-101` (definition: `cpu/quotients.rs:18-146`).
use itertools::Itertools;
use stwo_prover::core::backend::Column;
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::pcs::quotients::{ColumnSampleBatch, QuotientOps};
use stwo_prover::core::poly::circle::{CircleDomain, CircleEvaluation, SecureEvaluation};
use stwo_prover::core::poly::BitReversedOrder;
use stwo_prover::core::secure_column::SecureColumnByCoords;

use crate::backend::CudaBackend;
use crate::column::DeviceColumn;
use crate::ffi::*;
use crate::poly::point_words;

impl QuotientOps for CudaBackend {
    fn accumulate_quotients(
        domain: CircleDomain,
        columns: &[&CircleEvaluation<Self, BaseField, BitReversedOrder>],
        random_coeff: SecureField,
        sample_batches: &[ColumnSampleBatch],
        _log_blowup_factor: u32,
    ) -> SecureEvaluation<Self, BitReversedOrder> {
        // flatten the batches: points (8 words each), batch_start (prefix offsets into col_idx / values), per entry the
        // column index and the sampled value (4 words); batch order = first appearance of the point (IndexMap order,
        // `pcs/quotients.rs:50-69`), entry order = insertion order -- both feed the random-coefficient powers
        let cols = columns.iter().map(|c| c.values.as_ptr()).collect_vec();
        let points = sample_batches.iter().flat_map(|b| point_words(b.point)).collect_vec();
        let mut batch_start = vec![0u32];
        let (mut col_idx, mut values) = (Vec::new(), Vec::new());
        for b in sample_batches {
            for (idx, v) in &b.columns_and_values {
                col_idx.push(*idx as u32);
                values.extend(v.to_m31_array().map(|x| x.0));
            }
            batch_start.push(col_idx.len() as u32);
        }
        let alpha = random_coeff.to_m31_array().map(|x| x.0);
        let mut out: [DeviceColumn; 4] = std::array::from_fn(|_| unsafe { DeviceColumn::uninitialized(domain.size()) });
        let out_ptrs = out.iter_mut().map(|c| c.as_mut_ptr()).collect_vec();
        check(unsafe {
            cm31_accumulate_quotients(
                domain.log_size(), cols.as_ptr(), cols.len(), alpha.as_ptr(), sample_batches.len(), points.as_ptr(), batch_start.as_ptr(),
                col_idx.as_ptr(), values.as_ptr(), out_ptrs.as_ptr(),
            )
        });
        SecureEvaluation::new(domain, SecureColumnByCoords { columns: out })
    }
}
