//! `extern "C"` view of include/dplonk.h (the subset the worker needs; one line per symbol, same order as the header).
//! Every function returns 0 (DP_OK) or a negative DP_E_* code and never unwinds; `dp_last_error` gives the text.
#![allow(non_camel_case_types, dead_code)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct dp_ctx {
    _private: [u8; 0],
}

/// FftWorkload of utils.rs:3-9 as it crosses the ABI (include/dplonk.h: dp_fft_workload)
#[repr(C)]
#[derive(Clone, Copy)]
pub struct dp_fft_workload {
    pub row_start: u64,
    pub row_end: u64,
    pub col_start: u64,
    pub col_end: u64,
}

pub const DP_IPC_HANDLE_BYTES: usize = 64;

extern "C" {
    pub fn dp_create(cuda_device: c_int, me: u64, n_workers: u64, out: *mut *mut dp_ctx) -> c_int;
    pub fn dp_destroy(ctx: *mut dp_ctx) -> c_int;
    pub fn dp_last_error(ctx: *const dp_ctx) -> *const c_char;
    pub fn dp_init(ctx: *mut dp_ctx, bases: *const u8, n_bases: usize, domain_size: u64, quot_domain_size: u64) -> c_int;
    pub fn dp_msm(ctx: *mut dp_ctx, start: u64, end: u64, scalars: *const u8, n_scalars: usize, out144: *mut u8) -> c_int;
    pub fn dp_msm_submit(ctx: *mut dp_ctx, id: u64, start: u64, end: u64, scalars: *const u8, n_scalars: usize) -> c_int;
    pub fn dp_msm_collect(ctx: *mut dp_ctx, id: u64, out144: *mut u8) -> c_int;
    pub fn dp_commit(ctx: *mut dp_ctx, coeffs: *const u8, n: usize, out144: *mut u8) -> c_int;
    pub fn dp_fft_init(ctx: *mut dp_ctx, id: u64, workloads: *const dp_fft_workload, n_workloads: usize,
                       is_quot: c_int, is_inv: c_int, is_coset: c_int) -> c_int;
    pub fn dp_fft1(ctx: *mut dp_ctx, id: u64, i: u64, row: *const u8, len: usize) -> c_int;
    pub fn dp_fft1_rows(ctx: *mut dp_ctx, id: u64, i_first: u64, n_rows: u64, rows: *const u8) -> c_int;
    pub fn dp_fft1_rows_short(ctx: *mut dp_ctx, id: u64, i_first: u64, n_rows: u64, rows: *const u8, row_len: usize) -> c_int;
    pub fn dp_fft2_prepare(ctx: *mut dp_ctx, id: u64) -> c_int;
    pub fn dp_fft_exchange_begin(ctx: *mut dp_ctx, id: u64, send_dev: *mut *mut c_void, recv_dev: *mut *mut c_void,
                                 block_elems: *mut u64) -> c_int;
    pub fn dp_fft_exchange_begin_async(ctx: *mut dp_ctx, id: u64, send_dev: *mut *mut c_void, recv_dev: *mut *mut c_void,
                                       block_elems: *mut u64) -> c_int;
    pub fn dp_compute_stream(ctx: *mut dp_ctx, stream: *mut *mut c_void) -> c_int;
    pub fn dp_fft_exchange_end(ctx: *mut dp_ctx, id: u64) -> c_int;
    pub fn dp_fft2(ctx: *mut dp_ctx, id: u64, out: *mut u8, out_bytes: usize) -> c_int;
    pub fn dp_round1(ctx: *mut dp_ctx, evals: *const u8, n: usize, blind_2fr: *const u8, out144: *mut u8) -> c_int;
    pub fn dp_get_wire(ctx: *mut dp_ctx, out: *mut u8, out_bytes: usize, n_coeffs: *mut usize) -> c_int;
    pub fn dp_peer_arena_create(ctx: *mut dp_ctx, arena_bytes: u64, handle_out: *mut u8) -> c_int;
    pub fn dp_peer_attach(ctx: *mut dp_ctx, peer: u64, handle: *const u8) -> c_int;
    pub fn dp_peer_ready(ctx: *const dp_ctx) -> c_int;
    pub fn dp_sync(ctx: *mut dp_ctx) -> c_int;
}

/// non-zero return code -> the capnp error the caller of the RPC sees (the reference `unwrap()`s and panics instead)
pub fn check(ctx: *const dp_ctx, rc: c_int) -> Result<(), capnp::Error> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(dp_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(capnp::Error::failed(format!("dplonk error {}: {}", rc, msg)))
}
