//! FFI binding of `libmsm_b200.so` (C ABI: `include/msm_b200.h`) and a safe wrapper whose methods keep the names of
//! the reference's `SimulationObject` (`simulator/src/simulation_object.rs`).
//!
//! **Not compiled or tested where this repository is built (no Rust toolchain there).**  The tested hosts are the C++
//! layer `msm_b200/csrc/sim.cpp` and the Python/ctypes layer `msm_b200/`.
//!
//! Intended use inside `simulator/src/main.rs` (replacing `main.rs:43-79`):
//! ```ignore
//! let streams: Vec<SimulationParameters<f64>> = parameters_from_toml(toml).collect();
//! let mut sim = B200Simulation::new(&params_from(&streams[0], streams.len()))?;
//! for (s, p) in streams.iter().enumerate() { sim.set_psi(s, &build_ic_on_host(p))?; }
//! while sim.not_finished() {
//!     sim.update()?;                                   // panics upstream on Err(Aliasing), like :607-617
//!     for s in 0..streams.len() { let st = sim.state(s)?; if st.dumped == 1 { /* complex_array_to_disk */ } }
//! }
//! ```
use num::Complex;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

pub const MSM_OK: c_int = 0;
pub const MSM_E_ARG: c_int = -1;
pub const MSM_E_CUDA: c_int = -2;
pub const MSM_E_NCCL: c_int = -3;
pub const MSM_E_ALIASING: c_int = -4;
pub const MSM_E_NAN: c_int = -5;
pub const MSM_E_STATE: c_int = -6;
pub const MSM_E_NOMEM: c_int = -7;
pub const MSM_E_IO: c_int = -8;
pub const MSM_COUPLING_INDEPENDENT: i32 = 0;
pub const MSM_COUPLING_SUMMED: i32 = 1;

/// `msm_sim_params`: the resolved scalars of `SimulationParameters` (simulation_object.rs:67-140).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct MsmSimParams {
    pub struct_size: i32,
    pub dims: i32,
    pub size: i32,
    pub n_streams: i32,
    pub expanding: i32,
    pub coupling: i32,
    pub device: i32,
    pub chunk_streams: i32,
    pub num_data_dumps: u32,
    pub has_max_dloga: i32,
    pub rank: i32,
    pub nranks: i32,
    pub n_streams_global: i32,
    pub reserved: i32,
    pub axis_length: f64,
    pub time: f64,
    pub final_sim_time: f64,
    pub cfl: f64,
    pub total_mass: f64,
    pub particle_mass: f64,
    pub hbar_: f64,
    pub k2_cutoff: f64,
    pub alias_threshold: f64,
    pub omega_matter_now: f64,
    pub omega_radiation_now: f64,
    pub h: f64,
    pub z0: f64,
    pub max_dloga: f64,
    pub nccl_unique_id: *const c_void,
}

/// `msm_stream_state`.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct MsmStreamState {
    pub time: f64,
    pub tau: f64,
    pub dt: f64,
    pub potential_max: f64,
    pub alias_mass: f64,
    pub scale_factor: f64,
    pub n_steps: u64,
    pub current_dumps: u32,
    pub dumped: i32,
    pub finished: i32,
    pub aliased: i32,
}

#[repr(C)]
pub struct MsmSim {
    _private: [u8; 0],
}
#[repr(C)]
pub struct MsmCtx {
    _private: [u8; 0],
}

extern "C" {
    pub fn msm_version() -> *const c_char;
    pub fn msm_strerror(code: c_int) -> *const c_char;
    pub fn msm_nccl_unique_id(out128: *mut c_void) -> c_int;
    pub fn msm_sim_create(p: *const MsmSimParams, out: *mut *mut MsmSim) -> c_int;
    pub fn msm_sim_destroy(sim: *mut MsmSim);
    pub fn msm_sim_last_error(sim: *const MsmSim) -> *const c_char;
    pub fn msm_sim_ctx(sim: *mut MsmSim) -> *mut MsmCtx;
    pub fn msm_sim_set_psi(sim: *mut MsmSim, stream: i32, psi_interleaved: *const f64) -> c_int;
    pub fn msm_sim_update(sim: *mut MsmSim) -> c_int;
    pub fn msm_sim_update_streams(sim: *mut MsmSim, subset: *const i32) -> c_int;
    pub fn msm_sim_run_streams(sim: *mut MsmSim, n: i32, streams: *const i32, psi_in: *const *const f64,
                               re_out: *const *mut f64, im_out: *const *mut f64, max_updates: u64) -> c_int;
    pub fn msm_sim_not_finished(sim: *const MsmSim) -> c_int;
    pub fn msm_sim_state(sim: *const MsmSim, stream: i32, out: *mut MsmStreamState) -> c_int;
    pub fn msm_sim_get_psi(sim: *mut MsmSim, stream: i32, re: *mut f64, im: *mut f64) -> c_int;
    pub fn msm_sim_dump(sim: *mut MsmSim, stream: i32, root: *const c_char, name: *const c_char, idx: u32) -> c_int;
    pub fn msm_sim_dump_potential(sim: *mut MsmSim, stream: i32, root: *const c_char, name: *const c_char, idx: u32) -> c_int;
    pub fn msm_sim_wait_io(sim: *mut MsmSim) -> c_int;
    pub fn msm_sim_reserve_dump_buffers(sim: *mut MsmSim, n: i32) -> c_int;
    pub fn msm_sim_run_streams_seeded(sim: *mut MsmSim, n: i32, streams: *const i32, scheme: i32, seeds: *const u64,
                                      re_out: *const *mut f64, im_out: *const *mut f64, max_updates: u64) -> c_int;
    // grid level, for hosts that keep get_timestep / the scale-factor solver in Rust
    pub fn msm_potential_max(ctx: *mut MsmCtx, active: *const i32, max_abs_phi: *mut f64) -> c_int;
    pub fn msm_step(ctx: *mut MsmCtx, active: *const i32, drift: *const f64, kick: *const f64, alias: *mut f64) -> c_int;
    pub fn msm_get_potential(ctx: *mut MsmCtx, stream: i32, phi: *mut f64) -> c_int;
    pub fn msm_ensemble_accumulate(ctx: *mut MsmCtx, active: *const i32) -> c_int;
    pub fn msm_ensemble_allreduce(ctx: *mut MsmCtx) -> c_int;
    pub fn msm_ensemble_get(ctx: *mut MsmCtx, field: i32, re: *mut f64, im: *mut f64) -> c_int;
    pub fn msm_ic_store(ctx: *mut MsmCtx, stream: i32) -> c_int;
    pub fn msm_ic_load(ctx: *mut MsmCtx, stream: i32) -> c_int;
}

#[derive(Debug)]
pub enum RuntimeError {
    /// utils/error.rs `FourierAliasing`; the message names stream, threshold, k2_cutoff and p_mass
    FourierAliasing(String),
    Other(c_int, String),
}

/// All streams of one TOML, resident on one B200.
pub struct B200Simulation {
    raw: *mut MsmSim,
    cells: usize,
}

impl B200Simulation {
    pub fn new(params: &MsmSimParams) -> Result<Self, RuntimeError> {
        let mut p = *params;
        p.struct_size = std::mem::size_of::<MsmSimParams>() as i32;
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { msm_sim_create(&p, &mut raw) };
        if rc != MSM_OK {
            let msg = unsafe { CStr::from_ptr(msm_sim_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(RuntimeError::Other(rc, msg));
        }
        Ok(Self { raw, cells: (p.size as usize).pow(p.dims as u32) })
    }

    fn check(&self, rc: c_int) -> Result<(), RuntimeError> {
        if rc == MSM_OK {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(msm_sim_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(if rc == MSM_E_ALIASING { RuntimeError::FourierAliasing(msg) } else { RuntimeError::Other(rc, msg) })
    }

    /// `Array::new(&data, dim4)` of ics.rs:726: the host vector is dim-0-fastest, `Complex<f64>` is (re, im).
    pub fn set_psi(&mut self, stream: usize, psi: &[Complex<f64>]) -> Result<(), RuntimeError> {
        assert_eq!(psi.len(), self.cells);
        self.check(unsafe { msm_sim_set_psi(self.raw, stream as i32, psi.as_ptr() as *const f64) })
    }

    /// One `update()` (simulation_object.rs:475 / :669) for every unfinished stream.
    pub fn update(&mut self) -> Result<(), RuntimeError> {
        self.check(unsafe { msm_sim_update(self.raw) })
    }

    /// The stream loop of simulator/src/main.rs:43-85 in one call: upload `ics[i]` as stream i, `while not_finished()
    /// { update() }`, final psi into `re[i]` / `im[i]`.  Transfers of neighbouring stream groups overlap the step kernels
    /// (pin the vectors with cudaHostRegister for the overlap).
    pub fn run_streams(&mut self, ics: &[Vec<Complex<f64>>], re: &mut [Vec<f64>], im: &mut [Vec<f64>]) -> Result<(), RuntimeError> {
        let n = ics.len();
        assert!(re.len() == n && im.len() == n);
        let ids: Vec<i32> = (0..n as i32).collect();
        let ins: Vec<*const f64> = ics.iter().map(|v| { assert_eq!(v.len(), self.cells); v.as_ptr() as *const f64 }).collect();
        let rp: Vec<*mut f64> = re.iter_mut().map(|v| { v.resize(self.cells, 0.0); v.as_mut_ptr() }).collect();
        let ip: Vec<*mut f64> = im.iter_mut().map(|v| { v.resize(self.cells, 0.0); v.as_mut_ptr() }).collect();
        self.check(unsafe { msm_sim_run_streams(self.raw, n as i32, ids.as_ptr(), ins.as_ptr(), rp.as_ptr(), ip.as_ptr(), 0) })
    }

    /// simulation_object.rs:1226-1228, over all streams.
    pub fn not_finished(&self) -> bool {
        unsafe { msm_sim_not_finished(self.raw) == 1 }
    }

    pub fn state(&self, stream: usize) -> Result<MsmStreamState, RuntimeError> {
        let mut st = MsmStreamState::default();
        self.check(unsafe { msm_sim_state(self.raw, stream as i32, &mut st) })?;
        Ok(st)
    }

    /// `.host()` of utils/io.rs:46-47, already split into the two planes `complex_array_to_disk` writes.
    pub fn get_psi(&mut self, stream: usize) -> Result<(Vec<f64>, Vec<f64>), RuntimeError> {
        let (mut re, mut im) = (vec![0.0; self.cells], vec![0.0; self.cells]);
        self.check(unsafe { msm_sim_get_psi(self.raw, stream as i32, re.as_mut_ptr(), im.as_mut_ptr()) })?;
        Ok((re, im))
    }
}

impl Drop for B200Simulation {
    fn drop(&mut self) {
        unsafe { msm_sim_destroy(self.raw) }
    }
}
