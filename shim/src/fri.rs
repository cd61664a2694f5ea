//! `FriOps for CudaBackend` (external/stwo/crates/prover/src/core/fri.rs:92-139); replaces `simd/fri.rs:24-165`
//! (definitions: `fri.rs:1132-1189`, `cpu/fri.rs:29-85`).
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::fri::FriOps;
use stwo_prover::core::poly::circle::SecureEvaluation;
use stwo_prover::core::poly::line::LineEvaluation;
use stwo_prover::core::poly::twiddles::TwiddleTree;
use stwo_prover::core::poly::BitReversedOrder;
use stwo_prover::core::secure_column::SecureColumnByCoords;
use stwo_prover::core::backend::Column;

use crate::backend::CudaBackend;
use crate::column::DeviceColumn;
use crate::ffi::*;

fn coords(c: &SecureColumnByCoords<CudaBackend>) -> [*const u32; 4] {
    std::array::from_fn(|k| c.columns[k].as_ptr())
}
fn coords_mut(c: &mut SecureColumnByCoords<CudaBackend>) -> [*mut u32; 4] {
    std::array::from_fn(|k| c.columns[k].as_mut_ptr())
}
fn words(v: SecureField) -> [u32; 4] {
    v.to_m31_array().map(|x| x.0)
}

impl FriOps for CudaBackend {
    fn fold_line(eval: &LineEvaluation<Self>, alpha: SecureField, twiddles: &TwiddleTree<Self>) -> LineEvaluation<Self> {
        let log_size = eval.len().ilog2();
        assert!(log_size >= 1, "fold_line: fewer than two evaluations");
        let mut out = SecureColumnByCoords { columns: std::array::from_fn(|_| unsafe { DeviceColumn::uninitialized(eval.len() / 2) }) };
        let (src, dst, a) = (coords(&eval.values), coords_mut(&mut out), words(alpha));
        check(unsafe { cm31_fold_line(src.as_ptr(), log_size, a.as_ptr(), twiddles.itwiddles.0.raw, dst.as_ptr()) });
        LineEvaluation::new(eval.domain().double(), out)
    }

    fn fold_circle_into_line(dst: &mut LineEvaluation<Self>, src: &SecureEvaluation<Self, BitReversedOrder>, alpha: SecureField, twiddles: &TwiddleTree<Self>) {
        assert_eq!(src.len() >> 1, dst.len(), "fold_circle_into_line: src is not double the length of dst");
        let log_size = src.len().ilog2();
        let (s, d, a) = (coords(&src.values), coords_mut(&mut dst.values), words(alpha));
        check(unsafe { cm31_fold_circle_into_line(d.as_ptr(), s.as_ptr(), log_size, a.as_ptr(), twiddles.itwiddles.0.raw) });
    }

    fn decompose(eval: &SecureEvaluation<Self, BitReversedOrder>) -> (SecureEvaluation<Self, BitReversedOrder>, SecureField) {
        let log_size = eval.len().ilog2();
        let mut out = SecureColumnByCoords { columns: std::array::from_fn(|_| unsafe { DeviceColumn::uninitialized(eval.len()) }) };
        let (s, d) = (coords(&eval.values), coords_mut(&mut out));
        let mut lambda = [0u32; 4];
        check(unsafe { cm31_decompose(s.as_ptr(), log_size, d.as_ptr(), lambda.as_mut_ptr()) });
        let lambda = SecureField::from_m31_array(lambda.map(BaseField::from_u32_unchecked));
        (SecureEvaluation::new(eval.domain, out), lambda)
    }
}
