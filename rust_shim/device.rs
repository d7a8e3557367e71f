//! device.rs -- the device backend of `Decoder`: what replaces the six
//! `BufReader<zstd::Decoder<BufReader<IoSlice<R>>>>` readers (decoder/mod.rs:32, 218-226).
//!
//! UNBUILT (no Rust toolchain in the build image).  Shape of the change in nafcodec/src/decoder/mod.rs:
//!
//!   * `setup_block!` keeps parsing `(original_size, compressed_size)` (mod.rs:212-213) but, under
//!     `#[cfg(feature = "cuda")]`, instead of building a zstd reader it reads the `compressed_size`
//!     bytes of every wanted section into a buffer and records them in a `nafgpu_archive`.
//!   * `Decoder` gains `device: Option<DeviceRecords>`; `next_record` (mod.rs:356-399) takes the
//!     record from it instead of pulling the six readers.  `mask_sequence` is not called: masking
//!     (including the quirk of mod.rs:413-416) already happened on the device.
//!   * The public surface (`DecoderBuilder`, `Decoder`, `Record`, `Header`, `Error`) is unchanged.
use std::borrow::Cow;
use std::ffi::CStr;

use super::ffi::*;
use crate::data::Record;
use crate::error::Error;

pub struct DeviceRecords {
    ctx: *mut nafgpu_ctx,
    result: nafgpu_result,
    next: u64,
    // keeps the compressed sections alive for the duration of nafgpu_decode only
    // `DecoderBuilder::buffer_size` (mod.rs:104-112) under the device backend: the decoded archive stays in device memory and
    // `result` holds records [window_first, window_first + result.n_records) only, about `window_bytes` of decoded data
    // (nafgpu_job_fetch_window); 0 = the whole result was copied back by `decode`.
    window_bytes: u64,
    window_first: u64,
    n_records: u64,
}

// `arc` feature (lib.rs:25-26): a context has no thread-local CUDA state (cudaSetDevice at every entry).
unsafe impl Send for DeviceRecords {}

fn to_error(ctx: *mut nafgpu_ctx, status: i32) -> Error {
    use std::io::{Error as IoError, ErrorKind};
    let msg = unsafe { CStr::from_ptr(nafgpu_last_error(ctx)) }.to_string_lossy().into_owned();
    match status {
        NAFGPU_ERR_UNEXPECTED_EOF => Error::Io(IoError::new(ErrorKind::UnexpectedEof, msg)),
        NAFGPU_ERR_INVALID_DATA | NAFGPU_ERR_UTF8 | NAFGPU_ERR_UNSUPPORTED => {
            Error::Io(IoError::new(ErrorKind::InvalidData, msg))
        }
        NAFGPU_ERR_NOMEM => Error::Io(IoError::new(ErrorKind::OutOfMemory, msg)),
        _ => Error::Io(IoError::new(ErrorKind::Other, msg)),
    }
}

impl DeviceRecords {
    pub fn decode(archive: &nafgpu_archive, want: u32, device: i32) -> Result<Self, Error> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { nafgpu_ctx_create(device, &mut ctx) };
        if rc != NAFGPU_OK {
            return Err(Error::Io(std::io::Error::new(
                std::io::ErrorKind::Other,
                unsafe { CStr::from_ptr(nafgpu_strerror(rc)) }.to_string_lossy().into_owned(),
            )));
        }
        let mut result: nafgpu_result = unsafe { std::mem::zeroed() };
        let rc = unsafe { nafgpu_decode(ctx, archive, want, &mut result) };
        if rc != NAFGPU_OK {
            let e = to_error(ctx, rc);
            unsafe { nafgpu_ctx_destroy(ctx) };
            return Err(e);
        }
        let n_records = result.n_records;
        Ok(Self { ctx, result, next: 0, window_bytes: 0, window_first: 0, n_records })
    }

    /// Same, with host memory bounded by `buffer_size`: prepare + run once, records fetched a window at a time.
    pub fn decode_windowed(archive: &nafgpu_archive, want: u32, device: i32, buffer_size: u64) -> Result<Self, Error> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { nafgpu_ctx_create(device, &mut ctx) };
        if rc != NAFGPU_OK {
            return Err(Error::Io(std::io::Error::new(
                std::io::ErrorKind::Other,
                unsafe { CStr::from_ptr(nafgpu_strerror(rc)) }.to_string_lossy().into_owned(),
            )));
        }
        let mut rc = unsafe { nafgpu_job_prepare(ctx, archive, 1, want) };
        if rc == NAFGPU_OK {
            rc = unsafe { nafgpu_job_run(ctx) };
        }
        if rc != NAFGPU_OK {
            let e = to_error(ctx, rc);
            unsafe { nafgpu_ctx_destroy(ctx) };
            return Err(e);
        }
        let mut me = Self {
            ctx,
            result: unsafe { std::mem::zeroed() },
            next: 0,
            window_bytes: buffer_size.max(1),
            window_first: 0,
            n_records: archive.header.number_of_sequences,
        };
        me.fetch_window(0)?;
        Ok(me)
    }

    fn fetch_window(&mut self, first: u64) -> Result<(), Error> {
        let rc = unsafe {
            nafgpu_job_fetch_window(self.ctx, 0, first, self.n_records - first, self.window_bytes, &mut self.result)
        };
        if rc != NAFGPU_OK {
            return Err(to_error(self.ctx, rc));
        }
        self.window_first = first;
        Ok(())
    }

    /// The body of `Decoder::next_record` for the device backend.
    pub fn next_record(&mut self) -> Result<Record<'static>, Error> {
        if self.window_bytes != 0 && self.next >= self.window_first + self.result.n_records {
            self.fetch_window(self.next)?;
        }
        let r = &self.result;
        let i = self.next - self.window_first;        // (index within the window; the whole archive is one window otherwise)
        self.next += 1;
        if r.record_status != 0 && i == r.first_bad_record {
            return Err(to_error(self.ctx, r.record_status));
        }
        let slice = |blob: *const u8, off: *const u64, strip: u64| -> Cow<'static, str> {
            let (a, b) = unsafe { (*off.add(i as usize), *off.add(i as usize + 1)) };
            let bytes = unsafe { std::slice::from_raw_parts(blob.add(a as usize), (b - a - strip) as usize) };
            // validated on the device (k_utf8_validate); copy out: Record<'static> owns its strings
            Cow::Owned(unsafe { String::from_utf8_unchecked(bytes.to_vec()) })
        };
        let id = (!r.ids.is_null() && i < r.n_ids).then(|| slice(r.ids, r.id_offsets, 1));
        let comment = (!r.comments.is_null() && i < r.n_comments).then(|| slice(r.comments, r.comment_offsets, 1));
        let has_len = !r.lengths.is_null() && i < r.n_lengths;
        let length = has_len.then(|| unsafe { *r.lengths.add(i as usize) });
        let sequence = (has_len && !r.sequence.is_null()).then(|| slice(r.sequence, r.record_offsets, 0));
        let quality = (has_len && !r.quality.is_null()).then(|| slice(r.quality, r.record_offsets, 0));
        Ok(Record { id, comment, sequence, quality, length })
    }
}

impl Drop for DeviceRecords {
    fn drop(&mut self) {
        unsafe { nafgpu_ctx_destroy(self.ctx) };
    }
}
