//! `MerkleOps<Blake2sMerkleHasher>` (external/stwo/crates/prover/src/core/vcs/ops.rs:25-45) and
//! `GrindOps<Blake2sChannel>` (`proof_of_work.rs:3-7`) for `CudaBackend`.
//!
//! Replaces `simd/blake2s.rs:60-142` and `simd/grind.rs:21-71`.  `commit_on_layer` alone is the drop-in; `commit_tree`
//! is the fused form `MerkleProver::commit` (`vcs/prover.rs:40-66`) should call instead of its per-layer loop: a layer
//! followed by column-free layers is one launch (`cm31_blake2s_commit_multi`), and everything at or below 2^7 (2^10 when
//! small columns are injected there) is one single-CTA launch (`cm31_blake2s_commit_top`).
use itertools::Itertools;
use stwo_prover::core::backend::{Col, Column};
use stwo_prover::core::channel::Blake2sChannel;
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::proof_of_work::GrindOps;
use stwo_prover::core::vcs::blake2_hash::Blake2sHash;
use stwo_prover::core::vcs::blake2_merkle::Blake2sMerkleHasher;
use stwo_prover::core::vcs::ops::MerkleOps;

use crate::backend::CudaBackend;
use crate::column::DeviceHashColumn;
use crate::ffi::*;
use crate::poly::col_ptrs;

impl MerkleOps<Blake2sMerkleHasher> for CudaBackend {
    fn commit_on_layer(
        log_size: u32,
        prev_layer: Option<&Col<Self, Blake2sHash>>,
        columns: &[&Col<Self, BaseField>],
    ) -> Col<Self, Blake2sHash> {
        if let Some(prev) = prev_layer {
            assert_eq!(prev.len(), 2 << log_size);
        }
        let mut out = unsafe { DeviceHashColumn::uninitialized(1 << log_size) };
        let cols = col_ptrs(columns);
        check(unsafe {
            cm31_blake2s_commit_layer(log_size, prev_layer.map_or(std::ptr::null(), |p| p.as_ptr()), cols.as_ptr(), cols.len(), out.as_mut_ptr())
        });
        out
    }
}

impl CudaBackend {
    /// The whole of `MerkleProver::commit` (`vcs/prover.rs:40-66`): columns sorted by size descending (stable), layers
    /// returned root first like `MerkleProver::layers`.
    pub fn commit_tree(columns: &[&Col<Self, BaseField>]) -> Vec<DeviceHashColumn> {
        if columns.is_empty() {
            return vec![<Self as MerkleOps<Blake2sMerkleHasher>>::commit_on_layer(0, None, &[])];
        }
        let sorted = columns.iter().copied().sorted_by_key(|c| std::cmp::Reverse(c.len())).collect_vec();
        let log = |c: &&Col<Self, BaseField>| c.len().ilog2();
        let max_log = log(&sorted[0]);
        let top_log = if sorted.iter().any(|c| (8..=10).contains(&log(c))) { 10 } else { 7 };
        let mut layers: Vec<DeviceHashColumn> = Vec::new(); // leaf side first
        let mut pos = 0;
        let mut log_size = max_log as i64;
        while log_size > top_log {
            let start = pos;
            while pos < sorted.len() && log(&sorted[pos]) as i64 == log_size {
                pos += 1;
            }
            let next_with_cols = sorted.get(pos).map_or(-1, |c| log(c) as i64);
            let mut n_levels = 1;
            while log_size - n_levels > top_log && log_size - n_levels > next_with_cols && n_levels < 9 {
                n_levels += 1;
            }
            let cols = col_ptrs(&sorted[start..pos]);
            let prev = layers.last().map_or(std::ptr::null(), |p| p.as_ptr());
            let mut outs = (0..n_levels).map(|l| unsafe { DeviceHashColumn::uninitialized(1 << (log_size - l)) }).collect_vec();
            let out_ptrs = outs.iter_mut().map(|o| o.as_mut_ptr()).collect_vec();
            check(unsafe {
                if n_levels == 1 {
                    cm31_blake2s_commit_layer(log_size as u32, prev, cols.as_ptr(), cols.len(), out_ptrs[0])
                } else {
                    cm31_blake2s_commit_multi(log_size as u32, prev, cols.as_ptr(), cols.len(), n_levels as u32, out_ptrs.as_ptr())
                }
            });
            layers.extend(outs);
            log_size -= n_levels;
        }
        // single-CTA top: columns of every remaining size are injected at their layer
        let top = log_size.max(0) as u32;
        let mut col_start = vec![0u32; top as usize + 2];
        let mut top_cols: Vec<*const u32> = Vec::new();
        for l in 0..=top {
            col_start[l as usize] = top_cols.len() as u32;
            top_cols.extend(sorted[pos..].iter().filter(|c| log(c) == l).map(|c| c.as_ptr()));
        }
        col_start[top as usize + 1] = top_cols.len() as u32;
        let mut outs = (0..=top).map(|l| unsafe { DeviceHashColumn::uninitialized(1 << l) }).collect_vec();
        let out_ptrs = outs.iter_mut().map(|o| o.as_mut_ptr()).collect_vec();
        let prev = layers.last().map_or(std::ptr::null(), |p| p.as_ptr());
        check(unsafe { cm31_blake2s_commit_top(top, prev, top_cols.as_ptr(), col_start.as_ptr(), out_ptrs.as_ptr()) });
        layers.extend(outs.into_iter().rev());
        layers.reverse();
        layers
    }
}

impl GrindOps<Blake2sChannel> for CudaBackend {
    /// The MINIMUM nonce, like `simd/grind.rs:35-38` (`find_map_first`) and `cpu/grind.rs:7-15`.
    fn grind(channel: &Blake2sChannel, pow_bits: u32) -> u64 {
        let digest = channel.digest();
        let mut nonce = 0u64;
        check(unsafe { cm31_grind_blake2s(digest.0.as_ptr(), pow_bits, &mut nonce) });
        nonce
    }
}
