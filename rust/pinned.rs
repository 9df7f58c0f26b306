//! Pinned (page-locked) host memory for the CUDA backend: the pinned analogue of compu_malloc / compu_free
//! (src/mem.rs:27-49) and of `Buffer<N>` (src/buffer.rs:4-49). UNCOMPILED in this image (no rustc).
//!
//! The batched entry points DMA straight from / to these buffers; ordinary slices work too but are staged by the driver.

use core::{ptr, slice};

use crate::cuda_sys as sys;

///Page-locked byte buffer with a cursor, the pinned counterpart of `Buffer<N>`.
pub struct PinnedBuffer {
    ptr: ptr::NonNull<u8>,
    cap: usize,
    cursor: usize,
}

impl PinnedBuffer {
    ///Returns `None` when no device is usable or the allocation fails
    pub fn new(cap: usize) -> Option<Self> {
        let p = unsafe { sys::cz_host_alloc(cap) } as *mut u8;
        ptr::NonNull::new(p).map(|ptr| Self { ptr, cap, cursor: 0 })
    }
    #[inline(always)]
    pub fn data(&self) -> &[u8] {
        unsafe { slice::from_raw_parts(self.ptr.as_ptr(), self.cursor) }
    }
    #[inline(always)]
    pub fn spare_capacity_mut(&mut self) -> &mut [u8] {
        unsafe { slice::from_raw_parts_mut(self.ptr.as_ptr().add(self.cursor), self.cap - self.cursor) }
    }
    #[inline(always)]
    pub fn advance(&mut self, n: usize) {
        debug_assert!(self.cursor + n <= self.cap);
        self.cursor += n;
    }
    #[inline(always)]
    pub fn consume(&mut self) {
        self.cursor = 0;
    }
}

impl Drop for PinnedBuffer {
    fn drop(&mut self) {
        unsafe { sys::cz_host_free(self.ptr.as_ptr() as *mut _) }
    }
}
