//! Safe wrapper mirroring `oxide_control::physics::Physics` (reference src/physics.rs) over libox_b200.so,
//! plus `BatchedPhysics`. NOT COMPILED in the build environment (no Rust toolchain); the executable mirrors are
//! include/ox_b200.hpp (C++) and oxide_control_b200/physics.py (Python).
use ox_b200_sys as sys;
use std::ffi::{CStr, CString};

/// Mirror of `oxide_control::error::Error` (src/error.rs:3-15) plus a CUDA variant.
#[derive(Debug)]
pub enum Error { Mujoco(String), Mjs(String), NameNotFound(&'static str), PhysicsDiverged, Cuda(String), Invalid(String) }

fn last_error() -> String { unsafe { CStr::from_ptr(sys::ox_last_error_message()).to_string_lossy().into_owned() } }
fn check(s: sys::ox_status) -> Result<(), Error> {
    match s {
        sys::OX_OK => Ok(()),
        sys::OX_ERR_PARSE | sys::OX_ERR_IO => Err(Error::Mujoco(last_error())),   // From<MjError>, src/error.rs:17-21
        sys::OX_ERR_COMPILE => Err(Error::Mjs(last_error())),                      // src/physics.rs:21
        sys::OX_ERR_CUDA => Err(Error::Cuda(last_error())),
        _ => Err(Error::Invalid(last_error())),
    }
}

pub struct Model { raw: *mut sys::ox_model }
impl Model {
    pub fn from_xml_string(xml: impl Into<String>) -> Result<Self, Error> {
        let c = CString::new(xml.into()).map_err(|e| Error::Mujoco(e.to_string()))?;
        let mut raw = std::ptr::null_mut();
        check(unsafe { sys::ox_model_from_xml_string(c.as_ptr(), &mut raw) })?;
        Ok(Self { raw })
    }
    pub fn from_xml(path: impl AsRef<std::path::Path>) -> Result<Self, Error> {
        let c = CString::new(path.as_ref().to_str().unwrap()).unwrap();            // same unwrap as src/physics.rs:13
        let mut raw = std::ptr::null_mut();
        check(unsafe { sys::ox_model_from_xml_path(c.as_ptr(), &mut raw) })?;
        Ok(Self { raw })
    }
    pub fn size(&self, name: &str) -> i32 { let c = CString::new(name).unwrap(); unsafe { sys::ox_model_size(self.raw, c.as_ptr()) } }
    pub fn object_id(&self, objtype: i32, name: &str) -> Option<i32> {
        let c = CString::new(name).ok()?;
        let i = unsafe { sys::ox_model_name2id(self.raw, objtype, c.as_ptr()) };
        if i < 0 { None } else { Some(i) }
    }
}
impl Drop for Model { fn drop(&mut self) { unsafe { sys::ox_model_free(self.raw) } } }
unsafe impl Send for Model {}
unsafe impl Sync for Model {}   // immutable after construction

/// nenv copies of mjData on one GPU; `&mut self` on every mutator = one stream, externally serialised.
pub struct BatchedPhysics { raw: *mut sys::ox_batch, nenv: usize }
impl BatchedPhysics {
    pub fn new(model: &Model, nenv: usize, f64_validation: bool, device: i32, env_id_offset: i64) -> Result<Self, Error> {
        let mut cfg = unsafe { std::mem::zeroed::<sys::ox_batch_config>() };
        unsafe { sys::ox_batch_config_default(&mut cfg) };
        cfg.nenv = nenv as i32; cfg.device = device; cfg.env_id_offset = env_id_offset;
        cfg.precision = if f64_validation { sys::OX_F64 } else { sys::OX_F32 };
        let mut raw = std::ptr::null_mut();
        check(unsafe { sys::ox_batch_create(model.raw, &cfg, &mut raw) })?;
        Ok(Self { raw, nenv })
    }
    pub fn nenv(&self) -> usize { self.nenv }
    pub fn step(&mut self, nsteps: i32) { let _ = unsafe { sys::ox_batch_step(self.raw, nsteps) }; }     // infallible like src/physics.rs:44
    pub fn forward(&mut self) { let _ = unsafe { sys::ox_batch_forward(self.raw) }; }
    /// controls in, one step, qpos / qvel out in one call (Action::apply + step + Observation::generate, src/lib.rs:63-66)
    pub fn step_io(&mut self, ctrl: &[f32], qpos: &mut [f32], qvel: &mut [f32]) -> Result<(), Error> {
        check(unsafe { sys::ox_batch_step_io(self.raw, ctrl.as_ptr() as *const _, qpos.as_mut_ptr() as *mut _, qvel.as_mut_ptr() as *mut _, sys::OX_F32, sys::OX_MEM_HOST) })
    }
    pub fn reset(&mut self, mask: Option<&[u8]>) { let _ = unsafe { sys::ox_batch_reset(self.raw, mask.map_or(std::ptr::null(), |m| m.as_ptr())) }; }
    pub fn sync(&mut self) -> Result<(), Error> { check(unsafe { sys::ox_batch_sync(self.raw) }) }
    /// bulk upload of controls, `[nenv][nu]` f32 from (ideally pinned) host memory
    pub fn set_ctrl(&mut self, ctrl: &[f32]) -> Result<(), Error> {
        check(unsafe { sys::ox_batch_set(self.raw, sys::OX_F_CTRL, ctrl.as_ptr() as *const _, sys::OX_F32, sys::OX_MEM_HOST, sys::OX_LAYOUT_ENV_MAJOR) })
    }
    pub fn get(&mut self, field: i32, out: &mut [f32]) -> Result<(), Error> {
        check(unsafe { sys::ox_batch_get(self.raw, field, out.as_mut_ptr() as *mut _, sys::OX_F32, sys::OX_MEM_HOST, sys::OX_LAYOUT_ENV_MAJOR) })
    }
    pub fn get1(&mut self, field: i32, env: usize, offset: usize, out: &mut [f64]) -> Result<Option<()>, Error> {
        match unsafe { sys::ox_batch_get1(self.raw, field, env as i32, offset as i32, out.len() as i32, out.as_mut_ptr()) } {
            sys::OX_ABSENT => Ok(None), s => check(s).map(Some),
        }
    }
    pub fn set1(&mut self, field: i32, env: usize, offset: usize, v: &[f64]) -> Result<Option<()>, Error> {
        match unsafe { sys::ox_batch_set1(self.raw, field, env as i32, offset as i32, v.len() as i32, v.as_ptr()) } {
            sys::OX_ABSENT => Ok(None), s => check(s).map(Some),
        }
    }
}
impl Drop for BatchedPhysics { fn drop(&mut self) { unsafe { sys::ox_batch_free(self.raw) } } }
unsafe impl Send for BatchedPhysics {}

/// One environment: same method set as `oxide_control::physics::Physics` (src/physics.rs:6-171).
pub struct Physics { model: Model, data: BatchedPhysics }
impl Physics {
    pub fn from_xml_string(xml: impl Into<String>) -> Result<Self, Error> {
        let model = Model::from_xml_string(xml)?;
        let data = BatchedPhysics::new(&model, 1, true, 0, 0)?;
        Ok(Self { model, data })
    }
    pub fn from_xml(p: impl AsRef<std::path::Path>) -> Result<Self, Error> {
        let model = Model::from_xml(p)?;
        let data = BatchedPhysics::new(&model, 1, true, 0, 0)?;
        Ok(Self { model, data })
    }
    pub fn model(&self) -> &Model { &self.model }
    pub fn step(&mut self) { self.data.step(1); let _ = self.data.sync(); }
    pub fn forward(&mut self) { self.data.forward(); let _ = self.data.sync(); }
    pub fn reset(&mut self) { self.data.reset(None); let _ = self.data.sync(); }
    fn get(&mut self, f: i32, off: usize) -> f64 { let mut v = [0.0]; let _ = self.data.get1(f, 0, off, &mut v); v[0] }
    pub fn time(&mut self) -> f64 { self.get(sys::OX_F_TIME, 0) }
    pub fn set_time(&mut self, t: f64) { let _ = self.data.set1(sys::OX_F_TIME, 0, 0, &[t]); }
    pub fn ctrl(&mut self, id: usize) -> f64 { self.get(sys::OX_F_CTRL, id) }
    pub fn set_ctrl(&mut self, id: usize, v: f64) { let _ = self.data.set1(sys::OX_F_CTRL, 0, id, &[v]); }
    pub fn act(&mut self, _id: usize) -> Option<f64> { None }                       // src/physics.rs:96-98: stateless
    pub fn qacc_warmstart(&mut self, dof: usize) -> f64 { self.get(sys::OX_F_QACC_WARMSTART, dof) }
    pub fn qfrc_applied(&mut self, dof: usize) -> f64 { self.get(sys::OX_F_QFRC_APPLIED, dof) }
    pub fn xfrc_applied(&mut self, body: usize) -> [f64; 6] { let mut v = [0.0; 6]; let _ = self.data.get1(sys::OX_F_XFRC_APPLIED, 0, 6 * body, &mut v); v }
    pub fn mocap_pos(&mut self, _body: usize) -> Option<[f64; 3]> { None }          // src/physics.rs:155-157
}

// ---- N1: Environment<T: Task> for a whole batch, evaluated on device (reference src/lib.rs:28-88) ----
/// `enum TimeStep<O>` for every env: `finished[e]` selects `Finish` (no discount) or `Step`.
pub struct BatchedTimeStep<'a> { pub observation: &'a [f32], pub reward: &'a [f32], pub discount: &'a [f32], pub finished: &'a [u8] }
pub struct BatchedEnvironment<'p> {
    raw: *mut sys::ox_env, physics: &'p mut BatchedPhysics, obs_dim: usize,
    obs: Vec<f32>, reward: Vec<f32>, discount: Vec<f32>, finished: Vec<u8>,
}
impl<'p> BatchedEnvironment<'p> {
    /// `Environment::new(physics, task)` (src/lib.rs:33-36); the task is declarative because it runs inside a kernel.
    pub fn new(physics: &'p mut BatchedPhysics, task: &sys::ox_task_spec) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { sys::ox_env_create(physics.raw, task, &mut raw) })?;
        let (n, d) = (physics.nenv(), unsafe { sys::ox_env_obs_dim(raw) } as usize);
        Ok(Self { raw, physics, obs_dim: d, obs: vec![0.0; n * d], reward: vec![0.0; n], discount: vec![0.0; n], finished: vec![0; n] })
    }
    pub fn physics_mut(&mut self) -> &mut BatchedPhysics { self.physics }                                   // src/lib.rs:45-47
    pub fn obs_dim(&self) -> usize { self.obs_dim }
    /// `Environment::reset` (src/lib.rs:62-65)
    pub fn reset(&mut self) -> Result<&[f32], Error> {
        check(unsafe { sys::ox_env_reset(self.raw, self.obs.as_mut_ptr() as *mut _, sys::OX_F32, sys::OX_MEM_HOST) })?;
        Ok(&self.obs)
    }
    /// `Environment::step` (src/lib.rs:67-87): `action` is `[nenv][nu]`.
    pub fn step(&mut self, action: &[f32]) -> Result<BatchedTimeStep<'_>, Error> {
        check(unsafe { sys::ox_env_step(self.raw, action.as_ptr() as *const _, self.obs.as_mut_ptr() as *mut _, self.reward.as_mut_ptr() as *mut _,
                                        self.discount.as_mut_ptr() as *mut _, self.finished.as_mut_ptr(), sys::OX_F32, sys::OX_MEM_HOST) })?;
        Ok(BatchedTimeStep { observation: &self.obs, reward: &self.reward, discount: &self.discount, finished: &self.finished })
    }
}
impl Drop for BatchedEnvironment<'_> { fn drop(&mut self) { unsafe { sys::ox_env_free(self.raw) } } }
