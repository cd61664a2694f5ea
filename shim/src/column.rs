//! Device columns: `Col<CudaBackend, BaseField>`, `Col<CudaBackend, SecureField>`, `Col<CudaBackend, Blake2sHash>`
//! (`Column<T>`, external/stwo/crates/prover/src/core/backend/mod.rs:46-65).
//!
//! Layout: a base column is a plain `u32[len]` of canonical M31 values in HBM -- what
//! `simd/column.rs:26-30` keeps on the host; a secure column is four base columns (coordinate order a, b, c, d); a hash
//! column is `u32[8 * len]`, node i at words 8i..8i+8 (`vcs/blake2_hash.rs:8-10`).
use std::ffi::c_void;
use std::fmt;

use stwo_prover::core::backend::{Column, ColumnOps};
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::{SecureField, SECURE_EXTENSION_DEGREE};
use stwo_prover::core::vcs::blake2_hash::Blake2sHash;

use crate::backend::CudaBackend;
use crate::ffi::*;

/// An owned allocation of `words` u32 in device memory (stream-ordered pool behind `cm31_malloc`).
pub struct DeviceBuf {
    ptr: *mut u32,
    words: usize,
}
// One process drives one GPU from one thread (SURVEY.md §8b "Threading"); the buffers themselves are plain memory.
unsafe impl Send for DeviceBuf {}
unsafe impl Sync for DeviceBuf {}

impl DeviceBuf {
    pub fn uninit(words: usize) -> Self {
        let mut p: *mut c_void = std::ptr::null_mut();
        check(unsafe { cm31_malloc(&mut p, 4 * words.max(1)) });
        Self { ptr: p as *mut u32, words }
    }
    pub fn zeros(words: usize) -> Self {
        let b = Self::uninit(words);
        check(unsafe { cm31_memset0(b.ptr as *mut c_void, 4 * words.max(1)) });
        b
    }
    pub fn from_host(src: &[u32]) -> Self {
        let b = Self::uninit(src.len());
        if !src.is_empty() {
            check(unsafe { cm31_h2d(b.ptr as *mut c_void, src.as_ptr() as *const c_void, 4 * src.len()) });
        }
        b
    }
    pub fn to_host(&self) -> Vec<u32> {
        let mut out = vec![0u32; self.words];
        if self.words != 0 {
            check(unsafe { cm31_d2h(out.as_mut_ptr() as *mut c_void, self.ptr as *const c_void, 4 * self.words) });
        }
        out
    }
    pub fn word(&self, i: usize) -> u32 {
        assert!(i < self.words);
        let mut w = 0u32;
        check(unsafe { cm31_d2h(&mut w as *mut u32 as *mut c_void, self.ptr.add(i) as *const c_void, 4) });
        w
    }
    pub fn set_word(&mut self, i: usize, w: u32) {
        assert!(i < self.words);
        check(unsafe { cm31_h2d(self.ptr.add(i) as *mut c_void, &w as *const u32 as *const c_void, 4) });
    }
    pub fn as_ptr(&self) -> *const u32 {
        self.ptr
    }
    pub fn as_mut_ptr(&mut self) -> *mut u32 {
        self.ptr
    }
    pub fn words(&self) -> usize {
        self.words
    }
}
impl Clone for DeviceBuf {
    fn clone(&self) -> Self {
        let b = Self::uninit(self.words);
        if self.words != 0 {
            check(unsafe { cm31_d2d(b.ptr as *mut c_void, self.ptr as *const c_void, 4 * self.words) });
        }
        b
    }
}
impl Drop for DeviceBuf {
    fn drop(&mut self) {
        unsafe { cm31_free(self.ptr as *mut c_void) };
    }
}

// ------------------------------------------------------------------ BaseField
#[derive(Clone)]
pub struct DeviceColumn {
    pub buf: DeviceBuf,
}
impl fmt::Debug for DeviceColumn {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "DeviceColumn(len = {})", self.buf.words())
    }
}
impl DeviceColumn {
    pub fn as_ptr(&self) -> *const u32 {
        self.buf.as_ptr()
    }
    pub fn as_mut_ptr(&mut self) -> *mut u32 {
        self.buf.as_mut_ptr()
    }
    pub fn from_u32s(words: &[u32]) -> Self {
        Self { buf: DeviceBuf::from_host(words) }
    }
}
impl Column<BaseField> for DeviceColumn {
    fn zeros(len: usize) -> Self {
        Self { buf: DeviceBuf::zeros(len) }
    }
    unsafe fn uninitialized(len: usize) -> Self {
        Self { buf: DeviceBuf::uninit(len) }
    }
    fn to_cpu(&self) -> Vec<BaseField> {
        // canonical on the device, so the unchecked constructor is exact
        self.buf.to_host().into_iter().map(BaseField::from_u32_unchecked).collect()
    }
    fn len(&self) -> usize {
        self.buf.words()
    }
    /// One device round trip: the generic decommitment code calls this per query per column;
    /// `prover::GatherQueue` batches those reads instead.
    fn at(&self, index: usize) -> BaseField {
        BaseField::from_u32_unchecked(self.buf.word(index))
    }
    fn set(&mut self, index: usize, value: BaseField) {
        self.buf.set_word(index, value.0);
    }
}
impl FromIterator<BaseField> for DeviceColumn {
    fn from_iter<I: IntoIterator<Item = BaseField>>(iter: I) -> Self {
        let host: Vec<u32> = iter.into_iter().map(|v| v.0).collect();
        Self::from_u32s(&host)
    }
}
impl ColumnOps<BaseField> for CudaBackend {
    type Column = DeviceColumn;
    fn bit_reverse_column(column: &mut Self::Column) {
        let n = column.len();
        assert!(n.is_power_of_two());
        check(unsafe { cm31_bit_reverse(column.as_mut_ptr(), n.ilog2()) });
    }
}

// ------------------------------------------------------------------ SecureField (only used as scratch by generic code)
#[derive(Clone, Debug)]
pub struct DeviceSecureColumn {
    pub coords: [DeviceColumn; SECURE_EXTENSION_DEGREE],
}
impl Column<SecureField> for DeviceSecureColumn {
    fn zeros(len: usize) -> Self {
        Self { coords: std::array::from_fn(|_| DeviceColumn::zeros(len)) }
    }
    unsafe fn uninitialized(len: usize) -> Self {
        Self { coords: std::array::from_fn(|_| DeviceColumn::uninitialized(len)) }
    }
    fn to_cpu(&self) -> Vec<SecureField> {
        let c: Vec<Vec<BaseField>> = self.coords.iter().map(|c| c.to_cpu()).collect();
        (0..self.len()).map(|i| SecureField::from_m31_array(std::array::from_fn(|k| c[k][i]))).collect()
    }
    fn len(&self) -> usize {
        self.coords[0].len()
    }
    fn at(&self, index: usize) -> SecureField {
        SecureField::from_m31_array(std::array::from_fn(|k| self.coords[k].at(index)))
    }
    fn set(&mut self, index: usize, value: SecureField) {
        let v = value.to_m31_array();
        for k in 0..SECURE_EXTENSION_DEGREE {
            self.coords[k].set(index, v[k]);
        }
    }
}
impl FromIterator<SecureField> for DeviceSecureColumn {
    fn from_iter<I: IntoIterator<Item = SecureField>>(iter: I) -> Self {
        let host: Vec<SecureField> = iter.into_iter().collect();
        Self { coords: std::array::from_fn(|k| host.iter().map(|v| v.to_m31_array()[k]).collect()) }
    }
}
impl ColumnOps<SecureField> for CudaBackend {
    type Column = DeviceSecureColumn;
    fn bit_reverse_column(column: &mut Self::Column) {
        for c in column.coords.iter_mut() {
            <CudaBackend as ColumnOps<BaseField>>::bit_reverse_column(c);
        }
    }
}

// ------------------------------------------------------------------ Blake2sHash
#[derive(Clone)]
pub struct DeviceHashColumn {
    pub buf: DeviceBuf, // 8 words per node
}
impl fmt::Debug for DeviceHashColumn {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "DeviceHashColumn(len = {})", self.buf.words() / 8)
    }
}
impl DeviceHashColumn {
    pub fn as_ptr(&self) -> *const u32 {
        self.buf.as_ptr()
    }
    pub fn as_mut_ptr(&mut self) -> *mut u32 {
        self.buf.as_mut_ptr()
    }
}
impl Column<Blake2sHash> for DeviceHashColumn {
    fn zeros(len: usize) -> Self {
        Self { buf: DeviceBuf::zeros(8 * len) }
    }
    unsafe fn uninitialized(len: usize) -> Self {
        Self { buf: DeviceBuf::uninit(8 * len) }
    }
    fn to_cpu(&self) -> Vec<Blake2sHash> {
        bytemuck::cast_slice::<u32, [u8; 32]>(&self.buf.to_host()).iter().map(|b| Blake2sHash(*b)).collect()
    }
    fn len(&self) -> usize {
        self.buf.words() / 8
    }
    fn at(&self, index: usize) -> Blake2sHash {
        let mut h = [0u8; 32];
        check(unsafe { cm31_d2h(h.as_mut_ptr() as *mut c_void, self.buf.as_ptr().add(8 * index) as *const c_void, 32) });
        Blake2sHash(h)
    }
    fn set(&mut self, index: usize, value: Blake2sHash) {
        check(unsafe { cm31_h2d(self.buf.as_mut_ptr().add(8 * index) as *mut c_void, value.0.as_ptr() as *const c_void, 32) });
    }
}
impl FromIterator<Blake2sHash> for DeviceHashColumn {
    fn from_iter<I: IntoIterator<Item = Blake2sHash>>(iter: I) -> Self {
        let host: Vec<[u8; 32]> = iter.into_iter().map(|h| h.0).collect();
        Self { buf: DeviceBuf::from_host(bytemuck::cast_slice(&host)) }
    }
}
impl ColumnOps<Blake2sHash> for CudaBackend {
    type Column = DeviceHashColumn;
    fn bit_reverse_column(_column: &mut Self::Column) {
        unimplemented!("hash columns are never bit-reversed by the prover")
    }
}
