//! `CudaBackend`: the zero-sized backend type (external/stwo/crates/prover/src/core/backend/mod.rs:19-36).
use serde::{Deserialize, Serialize};
use stwo_prover::core::backend::{Backend, BackendForChannel};
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::lookups::gkr_prover::{GkrMultivariatePolyOracle, GkrOps, Layer};
use stwo_prover::core::lookups::mle::{Mle, MleOps};
use stwo_prover::core::lookups::utils::UnivariatePoly;
use stwo_prover::core::vcs::blake2_merkle::Blake2sMerkleChannel;

use crate::ffi::{check, cm31_lane, cm31_lanes_join, cm31_set_device};

#[derive(Copy, Clone, Debug, Default, Serialize, Deserialize)]
pub struct CudaBackend;

impl Backend for CudaBackend {}
impl BackendForChannel<Blake2sMerkleChannel> for CudaBackend {}

impl CudaBackend {
    /// One process drives one GPU: call once before the first op.
    pub fn set_device(device: i32) {
        check(unsafe { cm31_set_device(device) });
    }
    /// Stream lanes (include/cm31.h `cm31_lane`): per-component work of components with <= 2^12 rows is issued on the
    /// side lane so its latency-bound launches overlap the large components' kernels; joined before every commitment.
    pub fn lane_for(log_size: u32) {
        check(unsafe { cm31_lane(if log_size <= 12 { 1 } else { 0 }) });
    }
    pub fn lanes_join() {
        check(unsafe { cm31_lanes_join() });
    }
}

// GKR lookups are required by the `Backend` bound (`backend/mod.rs:19-31`) but never called by cairo-m (no `gkr` in
// crates/prover/src): stubs.
impl MleOps<BaseField> for CudaBackend {
    fn fix_first_variable(_mle: Mle<Self, BaseField>, _assignment: SecureField) -> Mle<Self, SecureField> {
        unimplemented!("GKR is not on cairo-m's proving path")
    }
}
impl MleOps<SecureField> for CudaBackend {
    fn fix_first_variable(_mle: Mle<Self, SecureField>, _assignment: SecureField) -> Mle<Self, SecureField> {
        unimplemented!("GKR is not on cairo-m's proving path")
    }
}
impl GkrOps for CudaBackend {
    fn gen_eq_evals(_y: &[SecureField], _v: SecureField) -> Mle<Self, SecureField> {
        unimplemented!("GKR is not on cairo-m's proving path")
    }
    fn next_layer(_layer: &Layer<Self>) -> Layer<Self> {
        unimplemented!("GKR is not on cairo-m's proving path")
    }
    fn sum_as_poly_in_first_variable(_h: &GkrMultivariatePolyOracle<'_, Self>, _claim: SecureField) -> UnivariatePoly<SecureField> {
        unimplemented!("GKR is not on cairo-m's proving path")
    }
}
