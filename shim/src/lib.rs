//! `zerocaf_b200` -- batch (slice) operations for dusk-zerocaf's hot path on a B200, over the C ABI of
//! `libzerocaf_b200.so`.  The per-element operator traits of the reference (`Add`, `Sub`, `Mul`, `Neg`, `Identity`,
//! `Square`, `Double` on `&T` and `T`; /root/reference/src/traits.rs:10-63) stay with the CPU backend -- a kernel launch per
//! element would be slower than the CPU -- and hot loops switch from `iter().zip().map(|(p, s)| p * s)` to the slice
//! functions of [`backend::cuda::Gpu`].
#![allow(non_snake_case)]
pub mod backend;
pub use backend::cuda::{EdwardsPoint, FieldElement, Gpu, MsmGenerators, RistrettoPoint, Scalar, ZcError};
