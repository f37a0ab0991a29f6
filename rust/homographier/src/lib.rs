//! Drop-in for the reference crate `homographier` (same module layout: `homographier::homographier`).
pub mod homographier;
