//! `PolyOps for CudaBackend` (external/stwo/crates/prover/src/core/poly/circle/ops.rs:13-69).
//!
//! Replaces `simd/circle.rs:132-297` (interpolate / evaluate / eval_at_point / precompute_twiddles) and `simd/fft/*`.
//! The batched forms (`interpolate_columns`, `evaluate_polynomials`) are overridden: the commitment scheme hands over
//! ~2 200 columns of which most have 16 rows, so every distinct log size is ONE launch.
use std::collections::BTreeMap;
use std::sync::Arc;

use itertools::Itertools;
use stwo_prover::core::backend::{Col, Column};
use stwo_prover::core::circle::{CirclePoint, Coset};
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::poly::circle::{CanonicCoset, CircleDomain, CircleEvaluation, CirclePoly, PolyOps};
use stwo_prover::core::poly::twiddles::TwiddleTree;
use stwo_prover::core::poly::BitReversedOrder;
use stwo_prover::core::ColumnVec;

use crate::backend::CudaBackend;
use crate::column::DeviceColumn;
use crate::ffi::*;

/// `PolyOps::Twiddles`: the device twiddle tree (both x and 1/x trees live behind one handle, so `twiddles` and
/// `itwiddles` of a `TwiddleTree` share it).
#[derive(Clone)]
pub struct DeviceTwiddles(pub Arc<TwiddleHandle>);
pub struct TwiddleHandle {
    pub raw: *mut Cm31Twiddles,
    pub log_size: u32,
}
unsafe impl Send for TwiddleHandle {}
unsafe impl Sync for TwiddleHandle {}
impl Drop for TwiddleHandle {
    fn drop(&mut self) {
        unsafe { cm31_twiddles_destroy(self.raw) };
    }
}

fn secure_words(v: SecureField) -> [u32; 4] {
    let a = v.to_m31_array();
    [a[0].0, a[1].0, a[2].0, a[3].0]
}
pub(crate) fn point_words(p: CirclePoint<SecureField>) -> [u32; 8] {
    let (x, y) = (secure_words(p.x), secure_words(p.y));
    [x[0], x[1], x[2], x[3], y[0], y[1], y[2], y[3]]
}

impl CudaBackend {
    /// All (polynomial, point) pairs of `CommitmentSchemeProver::prove_values` (`pcs/prover.rs:95-106`) in one launch.
    pub fn eval_at_points_batch(polys: &[&CirclePoly<Self>], points: &[CirclePoint<SecureField>], point_idx: &[u32]) -> Vec<SecureField> {
        assert_eq!(polys.len(), point_idx.len());
        let ptrs = polys.iter().map(|p| p.coeffs.as_ptr()).collect_vec();
        let logs = polys.iter().map(|p| p.log_size()).collect_vec();
        let pts = points.iter().flat_map(|p| point_words(*p)).collect_vec();
        let mut out = vec![0u32; 4 * polys.len()];
        check(unsafe {
            cm31_eval_at_point_batch(ptrs.as_ptr(), logs.as_ptr(), polys.len(), pts.as_ptr(), points.len(), point_idx.as_ptr(), out.as_mut_ptr())
        });
        out.chunks_exact(4)
            .map(|w| SecureField::from_m31_array(std::array::from_fn(|k| BaseField::from_u32_unchecked(w[k]))))
            .collect()
    }
}

impl PolyOps for CudaBackend {
    type Twiddles = DeviceTwiddles;

    fn interpolate(eval: CircleEvaluation<Self, BaseField, BitReversedOrder>, itwiddles: &TwiddleTree<Self>) -> CirclePoly<Self> {
        let log_size = eval.domain.log_size();
        let mut values = eval.values; // consumed, transformed in place
        let p = values.as_mut_ptr();
        check(unsafe { cm31_interpolate_batch(&p, 1, log_size, itwiddles.itwiddles.0.raw) });
        CirclePoly::new(values)
    }

    fn interpolate_columns(
        columns: impl IntoIterator<Item = CircleEvaluation<Self, BaseField, BitReversedOrder>>,
        twiddles: &TwiddleTree<Self>,
    ) -> Vec<CirclePoly<Self>> {
        let mut cols: Vec<(u32, DeviceColumn)> = columns.into_iter().map(|e| (e.domain.log_size(), e.values)).collect();
        let mut by_size: BTreeMap<u32, Vec<*mut u32>> = BTreeMap::new();
        for (log_size, c) in cols.iter_mut() {
            by_size.entry(*log_size).or_default().push(c.as_mut_ptr());
        }
        for (log_size, ptrs) in by_size {
            CudaBackend::lane_for(log_size);
            check(unsafe { cm31_interpolate_batch(ptrs.as_ptr(), ptrs.len(), log_size, twiddles.itwiddles.0.raw) });
        }
        cols.into_iter().map(|(_, c)| CirclePoly::new(c)).collect()
    }

    fn eval_at_point(poly: &CirclePoly<Self>, point: CirclePoint<SecureField>) -> SecureField {
        Self::eval_at_points_batch(&[poly], &[point], &[0])[0]
    }

    fn extend(poly: &CirclePoly<Self>, log_size: u32) -> CirclePoly<Self> {
        // not on cairo-m's path; the coefficient order is backend-private, so go through values
        assert!(log_size >= poly.log_size());
        let twiddles = Self::precompute_twiddles(CanonicCoset::new(log_size).circle_domain().half_coset);
        Self::interpolate(Self::evaluate(poly, CanonicCoset::new(log_size).circle_domain(), &twiddles), &twiddles)
    }

    fn evaluate(poly: &CirclePoly<Self>, domain: CircleDomain, twiddles: &TwiddleTree<Self>) -> CircleEvaluation<Self, BaseField, BitReversedOrder> {
        assert!(domain.is_canonic(), "libcm31 evaluates on canonic domains (every domain stwo's prover uses)");
        let mut out = unsafe { DeviceColumn::uninitialized(domain.size()) };
        let (src, dst) = (poly.coeffs.as_ptr(), out.as_mut_ptr());
        check(unsafe { cm31_evaluate_batch(&src, &dst, 1, poly.log_size(), domain.log_size(), twiddles.twiddles.0.raw) });
        CircleEvaluation::new(domain, out)
    }

    fn evaluate_polynomials(
        polynomials: &ColumnVec<CirclePoly<Self>>,
        log_blowup_factor: u32,
        twiddles: &TwiddleTree<Self>,
    ) -> Vec<CircleEvaluation<Self, BaseField, BitReversedOrder>> {
        let mut outs: Vec<DeviceColumn> =
            polynomials.iter().map(|p| unsafe { DeviceColumn::uninitialized(1 << (p.log_size() + log_blowup_factor)) }).collect();
        let mut by_size: BTreeMap<u32, (Vec<*const u32>, Vec<*mut u32>)> = BTreeMap::new();
        for (p, o) in polynomials.iter().zip(outs.iter_mut()) {
            let e = by_size.entry(p.log_size()).or_default();
            e.0.push(p.coeffs.as_ptr());
            e.1.push(o.as_mut_ptr());
        }
        for (log_size, (src, dst)) in by_size {
            CudaBackend::lane_for(log_size);
            check(unsafe { cm31_evaluate_batch(src.as_ptr(), dst.as_ptr(), src.len(), log_size, log_size + log_blowup_factor, twiddles.twiddles.0.raw) });
        }
        polynomials
            .iter()
            .zip(outs)
            .map(|(p, o)| CircleEvaluation::new(CanonicCoset::new(p.log_size() + log_blowup_factor).circle_domain(), o))
            .collect()
    }

    fn precompute_twiddles(coset: Coset) -> TwiddleTree<Self> {
        // `coset` is the half coset of the largest circle domain served (crates/prover/src/prover.rs:56-60); the device tree is
        // built for that canonic tower (stwo's prover only ever passes `CanonicCoset(..).half_coset()` or a doubling of it).
        let mut raw = std::ptr::null_mut();
        let log_size = coset.log_size() + 1;
        check(unsafe { cm31_twiddles_create(log_size, &mut raw) });
        let handle = DeviceTwiddles(Arc::new(TwiddleHandle { raw, log_size }));
        TwiddleTree { root_coset: coset, twiddles: handle.clone(), itwiddles: handle }
    }
}

/// Raw view used by the other op modules.
pub(crate) fn col_ptrs(cols: &[&Col<CudaBackend, BaseField>]) -> Vec<*const u32> {
    cols.iter().map(|c| c.as_ptr()).collect()
}
