//! `zlib-cuda` encoder interface: compu's `Interface` vtable over libcompu_b200.so (B200, sm_100a).
//! Same shape as src/encoder/zlib_ng.rs. UNCOMPILED in this image (no rustc) — see rust/README.md.

use core::ptr;

use super::zlib_common::{ZlibOptions, ZlibStrategy};
use super::{Encode, EncodeOp, EncodeStatus, Encoder, Interface};
use crate::cuda_sys as sys;

static ZLIB_CUDA: Interface = Interface::new(reset_fn, encode_fn, drop_fn);

impl Interface {
    ///Creates encoder with the CUDA (B200) interface
    ///
    ///Returns `None` if there is no usable sm_100 device or the options are invalid. There is no CPU fallback.
    pub fn zlib_cuda(opts: ZlibOptions) -> Option<Encoder> {
        let strategy = match opts.strategy {
            ZlibStrategy::Default => 0,
            ZlibStrategy::Filtered => 1,
            ZlibStrategy::HuffmanOnly => 2,
            ZlibStrategy::Rle => 3,
            ZlibStrategy::Fixed => 4,
        };
        // same argument order as deflateInit2_(level, Z_DEFLATED, windowBits, memLevel, strategy) (src/encoder/zlib_ng.rs:77-79)
        let state = unsafe { sys::cz_encoder_new(opts.compression as _, opts.mode as _, opts.mem_level as _, strategy) };
        ptr::NonNull::new(state as *mut u8).map(|state| unsafe { ZLIB_CUDA.encoder(state, [0; 2]) })
    }
}

#[inline]
unsafe fn encode_fn(state: ptr::NonNull<u8>, input: *const u8, input_remain: usize, output: *mut u8, output_remain: usize, op: EncodeOp) -> Encode {
    let op = match op {
        EncodeOp::Process => sys::CZ_OP_PROCESS,
        EncodeOp::Flush => sys::CZ_OP_FLUSH, // a sync-flush point, as in the reference (src/encoder/mod.rs:338)
        EncodeOp::Finish => sys::CZ_OP_FINISH,
    };
    let r = sys::cz_encode(state.as_ptr() as *mut _, input, input_remain, output, output_remain, op);
    Encode {
        input_remain: r.input_remain,
        output_remain: r.output_remain,
        // the library already applied the reference's status map (src/encoder/mod.rs:357-367)
        status: match r.status {
            sys::CZ_ENCODE_CONTINUE => EncodeStatus::Continue,
            sys::CZ_ENCODE_NEED_OUTPUT => EncodeStatus::NeedOutput,
            sys::CZ_ENCODE_FINISHED => EncodeStatus::Finished,
            _ => EncodeStatus::Error,
        },
    }
}

#[inline]
fn reset_fn(state: ptr::NonNull<u8>, _: [u8; 2]) -> Option<ptr::NonNull<u8>> {
    ptr::NonNull::new(unsafe { sys::cz_encoder_reset(state.as_ptr() as *mut _) } as *mut u8)
}

#[inline]
fn drop_fn(state: ptr::NonNull<u8>) {
    unsafe { sys::cz_encoder_free(state.as_ptr() as *mut _) }
}
