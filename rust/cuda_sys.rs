//! Raw bindings of include/compu_b200.h (the C ABI of libcompu_b200.so). UNCOMPILED in this image (no rustc).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Copy, Clone)]
pub struct cz_result {
    pub input_remain: usize,
    pub output_remain: usize,
    pub status: i32,
}

pub const CZ_DECODE_NEED_INPUT: i32 = 0;
pub const CZ_DECODE_NEED_OUTPUT: i32 = 1;
pub const CZ_DECODE_FINISHED: i32 = 2;
pub const CZ_DECODE_NEED_DICT: i32 = 3;

pub const CZ_ENCODE_CONTINUE: i32 = 0;
pub const CZ_ENCODE_NEED_OUTPUT: i32 = 1;
pub const CZ_ENCODE_FINISHED: i32 = 2;

pub const CZ_OP_PROCESS: c_int = 0;
pub const CZ_OP_FLUSH: c_int = 1;
pub const CZ_OP_FINISH: c_int = 2;

extern "C" {
    pub fn cz_device_count() -> c_int;
    pub fn cz_last_error() -> *const c_char;
    pub fn cz_host_alloc(bytes: usize) -> *mut c_void;
    pub fn cz_host_free(p: *mut c_void);

    pub fn cz_set_stream_device(device: c_int) -> c_int;
    pub fn cz_decoder_new(window_bits: c_int) -> *mut c_void;
    pub fn cz_decode(state: *mut c_void, input: *const u8, in_len: usize, output: *mut u8, out_len: usize) -> cz_result;
    pub fn cz_decoder_reset(state: *mut c_void) -> *mut c_void;
    pub fn cz_decoder_free(state: *mut c_void);
    pub fn cz_describe_error(code: i32) -> *const c_char;

    pub fn cz_encoder_new(level: c_int, window_bits: c_int, mem_level: c_int, strategy: c_int) -> *mut c_void;
    pub fn cz_encode(state: *mut c_void, input: *const u8, in_len: usize, output: *mut u8, out_len: usize, op: c_int) -> cz_result;
    pub fn cz_encoder_reset(state: *mut c_void) -> *mut c_void;
    pub fn cz_encoder_free(state: *mut c_void);

    pub fn cz_inflate_batch(n: usize, input: *const u8, in_offsets: *const u64, output: *mut u8, out_offsets: *const u64,
                            out_lens: *mut u64, statuses: *mut i32, in_consumed: *mut u64, window_bits: c_int, devices_mask: u32) -> c_int;
    pub fn cz_deflate_bound(len: u64, window_bits: c_int, segment_bytes: u64) -> u64;
    pub fn cz_deflate_batch(n: usize, input: *const u8, in_offsets: *const u64, output: *mut u8, out_offsets: *const u64,
                            out_lens: *mut u64, statuses: *mut i32, level: c_int, window_bits: c_int, strategy: c_int,
                            segment_bytes: u64, devices_mask: u32) -> c_int;
    pub fn cz_deflate_segmented(input: *const u8, len: u64, output: *mut u8, cap: u64, out_len: *mut u64, level: c_int,
                                window_bits: c_int, strategy: c_int, segment_bytes: u64, devices_mask: u32,
                                seg_index: *mut u64, seg_index_cap: u64, n_segments: *mut u64) -> c_int;
    pub fn cz_inflate_segmented(input: *const u8, len: u64, output: *mut u8, cap: u64, out_len: *mut u64, window_bits: c_int,
                                segment_bytes: u64, seg_index: *const u64, n_segments: u64, devices_mask: u32) -> c_int;
}
