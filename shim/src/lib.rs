//! `CudaBackend`: stwo's `Backend` / `BackendForChannel<Blake2sMerkleChannel>` over `libcm31.so` (include/cm31.h).
//!
//! One module per trait of the drop-in boundary:
//!
//! | module | stwo trait (file:line under external/stwo/crates/prover/src/core) |
//! |---|---|
//! | [`column`] | `Column<T>`, `ColumnOps<T>` (`backend/mod.rs:38-65`) |
//! | [`poly`] | `PolyOps` (`poly/circle/ops.rs:13-69`) |
//! | [`merkle`] | `MerkleOps<Blake2sMerkleHasher>` (`vcs/ops.rs:25-45`), `GrindOps<Blake2sChannel>` (`proof_of_work.rs:3-7`) |
//! | [`quotients`] | `QuotientOps` (`pcs/quotients.rs:22-35`) |
//! | [`fri`] | `FriOps` (`fri.rs:92-139`) |
//! | [`accumulation`] | `AccumulationOps` (`air/accumulation.rs:156-162`) |
//! | [`backend`] | `Backend`, `BackendForChannel` (`backend/mod.rs:19-36`), `GkrOps` / `MleOps` stubs |
//! | [`bytecode`] | `ExprEvaluator` (`constraint_framework/src/expr/evaluator.rs:63`) -> AIR bytecode of csrc/host/air_expr.hpp |
//! | [`component_prover`] | `ComponentProver<CudaBackend> for FrameworkComponent<E>` (`constraint_framework/src/component.rs:282-424`) |
//! | [`prover`] | `prove_cairo_m` made generic enough to run on `CudaBackend` (`crates/prover/src/prover.rs:23-147`) |
//!
//! Every method body is one or two `cm31_*` calls; the C++ class `csrc/host/cuda_backend.hpp::CudaBackend` is the same
//! binding written in C++ and is what the parity tests of this repository exercise.
#![allow(clippy::missing_safety_doc)]

pub mod accumulation;
pub mod backend;
pub mod bytecode;
pub mod column;
pub mod component_prover;
pub mod ffi;
pub mod fri;
pub mod merkle;
pub mod poly;
pub mod prover;
pub mod quotients;

pub use backend::CudaBackend;
pub use column::{DeviceColumn, DeviceHashColumn, DeviceSecureColumn};
