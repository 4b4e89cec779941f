//! Mirrors /root/reference/src/backend/mod.rs:9-16: one module per backend, selected by cargo feature.
pub mod cuda;
