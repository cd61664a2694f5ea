//! `AccumulationOps for CudaBackend` (external/stwo/crates/prover/src/core/air/accumulation.rs:156-162); replaces
//! `simd/accumulation.rs:11-39`.
use stwo_prover::core::air::accumulation::AccumulationOps;
use stwo_prover::core::backend::Column;
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::secure_column::SecureColumnByCoords;

use crate::backend::CudaBackend;
use crate::ffi::*;

impl AccumulationOps for CudaBackend {
    fn accumulate(column: &mut SecureColumnByCoords<Self>, other: &SecureColumnByCoords<Self>) {
        let n = column.columns[0].len();
        assert_eq!(n, other.columns[0].len());
        let dst: [*mut u32; 4] = std::array::from_fn(|k| column.columns[k].as_mut_ptr());
        let src: [*const u32; 4] = std::array::from_fn(|k| other.columns[k].as_ptr());
        check(unsafe { cm31_accumulate(dst.as_ptr(), src.as_ptr(), n) });
    }

    fn generate_secure_powers(felt: SecureField, n_powers: usize) -> Vec<SecureField> {
        let f = felt.to_m31_array().map(|x| x.0);
        let mut out = vec![0u32; 4 * n_powers];
        check(unsafe { cm31_secure_powers(f.as_ptr(), n_powers, out.as_mut_ptr()) });
        out.chunks_exact(4).map(|w| SecureField::from_m31_array(std::array::from_fn(|k| BaseField::from_u32_unchecked(w[k])))).collect()
    }
}
