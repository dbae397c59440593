//! `extern "C"` bindings to `libquadrs_gpu.so` (include/quadrs_gpu.h) and a safe `GpuChain` that
//! implements the reference's `Samples` trait (quadrs `src/samples.rs:11-28`).
//!
//! UNBUILT in this repository's image (no Rust toolchain); kept in sync with the header by hand.
#![allow(non_camel_case_types)]

use std::ffi::{c_char, c_int, c_void, CStr};
use std::ptr;

use anyhow::{anyhow, ensure, Error};
use num_complex::Complex;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct qd_source {
    pub kind: i32,
    pub format: i32,
    pub sample_rate: u64,
    pub data: *const c_void,
    pub n_bytes: u64,
    pub path: *const c_char,
    pub base_sample: u64,
    pub total_samples: u64,
    pub gen_seconds: f64,
    pub gen_cos: *const i64,
    pub gen_n_cos: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct qd_stage {
    pub kind: i32,
    pub reserved: i32,
    pub frequency: i64,
    pub decimate: u64,
    pub size: u64,
}

#[repr(C)]
pub struct qd_chain {
    _private: [u8; 0],
}

pub const QD_OK: c_int = 0;
pub const QD_E_SHORT_READ: c_int = 6;
pub const QD_E_WRITE_SHORT: c_int = 13;
pub const QD_SRC_FILE: i32 = 2;
pub const QD_SRC_GEN: i32 = 3;
pub const QD_STAGE_SHIFT: i32 = 1;
pub const QD_STAGE_LOWPASS: i32 = 2;
pub const QD_SPACE_HOST: c_int = 0;

extern "C" {
    pub fn qd_last_error() -> *const c_char;
    pub fn qd_chain_create(src: *const qd_source, stages: *const qd_stage, n_stages: usize, device: c_int,
                           out: *mut *mut qd_chain) -> c_int;
    /// One process, several GPUs: every sink call fans out over `devices` and gathers into the caller's one buffer.
    pub fn qd_chain_create_sharded(src: *const qd_source, stages: *const qd_stage, n_stages: usize,
                                   devices: *const c_int, n_dev: usize, out: *mut *mut qd_chain) -> c_int;
    pub fn qd_chain_n_devices(c: *const qd_chain, n_dev: *mut usize) -> c_int;
    pub fn qd_chain_destroy(c: *mut qd_chain);
    pub fn qd_chain_len(c: *const qd_chain, len: *mut u64) -> c_int;
    pub fn qd_chain_sample_rate(c: *const qd_chain, rate: *mut u64) -> c_int;
    pub fn qd_chain_read_at(c: *mut qd_chain, off: u64, buf: *mut Complex<f32>, n: usize, space: c_int,
                            got: *mut usize) -> c_int;
    pub fn qd_sparkfft_rows(c: *const qd_chain, width: usize, stride: u64, rows: *mut u64) -> c_int;
    pub fn qd_sparkfft(c: *mut qd_chain, width: usize, stride: u64, has_range: c_int, min: f32, max: f32,
                       first_row: u64, n_rows: u64, idx_out: *mut u8, mag_out: *mut f32, space: c_int,
                       rows_out: *mut u64) -> c_int;
    pub fn qd_freq_levels(c: *mut qd_chain, width: usize, stride: u64, levels: usize, first: u64, n: u64,
                          vals: *mut u8, space: c_int, total: *mut u64) -> c_int;
    pub fn qd_take_fft(c: *mut qd_chain, has_slice: c_int, start: u64, end: u64, width: usize, windowing: c_int,
                       output_len: usize, out: *mut f32, space: c_int) -> c_int;
    pub fn qd_write_file(c: *mut qd_chain, prefix: *const c_char, overwrite: c_int, name_out: *mut c_char,
                         name_cap: usize) -> c_int;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(qd_last_error()).to_string_lossy().into_owned() }
}

/// The reference's trait, restated so this crate stands alone (quadrs would `use crate::Samples`).
pub trait Samples: Sync + Send {
    fn len(&self) -> u64;
    fn sample_rate(&self) -> u64;
    fn read_at(&self, off: u64, buf: &mut [Complex<f32>]) -> usize;
}

/// A whole `from ... | shift ... | lowpass ...` graph evaluated on the GPU.
pub struct GpuChain {
    raw: *mut qd_chain,
}

// calls on one chain are serialised by a mutex inside the library
unsafe impl Send for GpuChain {}
unsafe impl Sync for GpuChain {}

impl GpuChain {
    pub fn new(src: &qd_source, stages: &[qd_stage], device: i32) -> Result<Self, Error> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { qd_chain_create(src, stages.as_ptr(), stages.len(), device, &mut raw) };
        // construction-time panics of the reference (shift.rs:20-24) arrive as codes 2/3: re-panic
        match rc {
            QD_OK => Ok(GpuChain { raw }),
            2 | 3 => panic!("{}", last_error()),
            _ => Err(anyhow!(last_error())),
        }
    }

    /// fft::spark_fft (src/fft.rs:12-69): prints the header and one row per window.
    pub fn spark_fft(&mut self, width: usize, stride: u64, min: Option<f32>, max: Option<f32>) -> Result<(), Error> {
        println!("sparkfft sample_rate={}", self.sample_rate());
        let mut rows = 0u64;
        ensure!(unsafe { qd_sparkfft_rows(self.raw, width, stride, &mut rows) } == QD_OK, last_error());
        let mut idx = vec![0u8; rows as usize * width];
        let mut got = 0u64;
        let rc = unsafe {
            qd_sparkfft(self.raw, width, stride, (min.is_some() || max.is_some()) as c_int, min.unwrap_or(0.08),
                        max.unwrap_or(1.0), 0, rows, idx.as_mut_ptr(), ptr::null_mut(), QD_SPACE_HOST, &mut got)
        };
        let glyphs: Vec<char> = " ▁▂▃▄▅▆▇█".chars().collect();
        for row in idx.chunks(width).take(got as usize) {
            println!("│{}│", row.iter().map(|&i| glyphs[i as usize]).collect::<String>());
        }
        match rc {
            QD_OK => Ok(()),
            QD_E_SHORT_READ => Err(anyhow!(last_error())), // read_exact_at's Err (samples.rs:20-25)
            _ => panic!("{}", last_error()),               // the reference panics here
        }
    }
}

impl Samples for GpuChain {
    fn len(&self) -> u64 {
        let mut v = 0;
        if unsafe { qd_chain_len(self.raw, &mut v) } != QD_OK {
            panic!("{}", last_error()); // filter.rs:46
        }
        v
    }

    fn sample_rate(&self) -> u64 {
        let mut v = 0;
        unsafe { qd_chain_sample_rate(self.raw, &mut v) };
        v
    }

    fn read_at(&self, off: u64, buf: &mut [Complex<f32>]) -> usize {
        let mut got = 0usize;
        let rc = unsafe { qd_chain_read_at(self.raw, off, buf.as_mut_ptr(), buf.len(), QD_SPACE_HOST, &mut got) };
        if rc != QD_OK {
            panic!("{}", last_error()); // samples.rs:74, filter.rs:76
        }
        got
    }
}

impl Drop for GpuChain {
    fn drop(&mut self) {
        unsafe { qd_chain_destroy(self.raw) }
    }
}
