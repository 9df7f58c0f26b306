//! `zlib-cuda` decoder interface: compu's `Interface` vtable over libcompu_b200.so (B200, sm_100a).
//! Same shape as src/decoder/zlib_ng.rs. UNCOMPILED in this image (no rustc) — see rust/README.md.

use core::ptr;

use super::zlib_common::ZlibMode;
use super::{Decode, DecodeError, DecodeStatus, Decoder, Interface};
use crate::cuda_sys as sys;

static ZLIB_CUDA: Interface = Interface::new(decode_fn, reset_fn, drop_fn, describe_error_fn);

impl Interface {
    ///Creates decoder with the CUDA (B200) interface
    ///
    ///Returns `None` if there is no usable sm_100 device or the state cannot be created. There is no CPU fallback.
    pub fn zlib_cuda(mode: ZlibMode) -> Option<Decoder> {
        // ZlibMode::max_bits(): Deflate -15, Zlib 15, Gzip 31, Auto 47 (src/decoder/zlib_common.rs:19-21)
        let state = unsafe { sys::cz_decoder_new(mode.max_bits()) };
        ptr::NonNull::new(state as *mut u8).map(|state| unsafe { ZLIB_CUDA.decoder(state) })
    }
}

#[inline]
unsafe fn decode_fn(state: ptr::NonNull<u8>, input: *const u8, input_remain: usize, output: *mut u8, output_remain: usize) -> Decode {
    let r = sys::cz_decode(state.as_ptr() as *mut _, input, input_remain, output, output_remain);
    Decode {
        input_remain: r.input_remain,
        output_remain: r.output_remain,
        // the library already applied the reference's status map (src/decoder/mod.rs:472-484)
        status: match r.status {
            sys::CZ_DECODE_NEED_INPUT => Ok(DecodeStatus::NeedInput),
            sys::CZ_DECODE_NEED_OUTPUT => Ok(DecodeStatus::NeedOutput),
            sys::CZ_DECODE_FINISHED => Ok(DecodeStatus::Finished),
            sys::CZ_DECODE_NEED_DICT => Err(DecodeError(2)), // Z_NEED_DICT
            code => Err(DecodeError(code)),
        },
    }
}

#[inline]
fn reset_fn(state: ptr::NonNull<u8>) -> Option<ptr::NonNull<u8>> {
    // returned pointer MUST replace the old one (src/decoder/mod.rs:433-441); this backend hands the same one back
    ptr::NonNull::new(unsafe { sys::cz_decoder_reset(state.as_ptr() as *mut _) } as *mut u8)
}

#[inline]
fn drop_fn(state: ptr::NonNull<u8>) {
    unsafe { sys::cz_decoder_free(state.as_ptr() as *mut _) }
}

#[inline]
fn describe_error_fn(code: i32) -> Option<&'static str> {
    // static storage, same text as zError (src/decoder/zlib_ng.rs:118-123)
    crate::utils::convert_c_str(unsafe { sys::cz_describe_error(code) } as *const i8)
}
