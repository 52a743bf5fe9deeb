//! Raw bindings to libox_b200.so — transcription of include/ox_b200.h.
//! NOT COMPILED in the build environment (no cargo/rustc there); kept in sync with the header by
//! tests/test_abi.py, which checks the header against the ctypes table this file mirrors.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

pub type ox_status = i32;
pub const OX_OK: ox_status = 0;
pub const OX_ERR_PARSE: ox_status = 1; // -> Error::Mujoco
pub const OX_ERR_COMPILE: ox_status = 2; // -> Error::Mjs
pub const OX_ERR_CUDA: ox_status = 3;
pub const OX_ERR_INVALID: ox_status = 4;
pub const OX_ABSENT: ox_status = 5; // -> Option::None
pub const OX_ERR_IO: ox_status = 6; // -> Error::Mujoco

#[repr(C)] pub struct ox_model { _private: [u8; 0] }
#[repr(C)] pub struct ox_batch { _private: [u8; 0] }
#[repr(C)] pub struct ox_env { _private: [u8; 0] }
#[repr(C)] #[derive(Clone, Copy)] pub struct ox_obs_segment { pub field: i32, pub first: i32, pub count: i32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct ox_reward_term { pub field: i32, pub index: i32, pub kind: i32, pub reserved_: i32, pub weight: c_double }
#[repr(C)] #[derive(Clone, Copy)] pub struct ox_finish_cond { pub field: i32, pub index: i32, pub lo: c_double, pub hi: c_double }
#[repr(C)]
#[derive(Clone, Copy)]
pub struct ox_task_spec {
    pub nobs: i32, pub obs: *const ox_obs_segment, pub nreward: i32, pub reward: *const ox_reward_term,
    pub nfinish: i32, pub finish: *const ox_finish_cond, pub reward_bias: c_double, pub time_limit: c_double,
    pub discount: c_double, pub init_qpos_noise: c_double, pub init_qvel_noise: c_double, pub seed: u64,
    pub frame_skip: i32, pub auto_reset: i32,
}
pub const OX_REWARD_LINEAR: i32 = 0; pub const OX_REWARD_SQUARE: i32 = 1; pub const OX_REWARD_ABS: i32 = 2;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ox_batch_config {
    pub nenv: i32, pub device: i32, pub precision: i32, pub mode: i32,
    pub iterations: i32, pub ls_iterations: i32, pub use_graph: i32, pub block_threads: i32,
    pub env_id_offset: i64, pub tolerance: c_double, pub specialize: i32, pub lanes_per_warp: i32, pub coop_solver: i32, pub reserved_: i32,
}

pub const OX_F32: i32 = 0; pub const OX_F64: i32 = 1;
pub const OX_MEM_HOST: i32 = 0; pub const OX_MEM_DEVICE: i32 = 1;
pub const OX_LAYOUT_ENV_MAJOR: i32 = 0; pub const OX_LAYOUT_ELEM_MAJOR: i32 = 1;
pub const OX_F_QPOS: i32 = 0; pub const OX_F_QVEL: i32 = 1; pub const OX_F_CTRL: i32 = 2; pub const OX_F_QFRC_APPLIED: i32 = 3;
pub const OX_F_XFRC_APPLIED: i32 = 4; pub const OX_F_QACC_WARMSTART: i32 = 5; pub const OX_F_TIME: i32 = 6; pub const OX_F_ACT: i32 = 7;
pub const OX_F_QACC: i32 = 8; pub const OX_F_SENSORDATA: i32 = 9; pub const OX_F_DIVERGED: i32 = 103;

#[link(name = "ox_b200")]
extern "C" {
    pub fn ox_last_error_message() -> *const c_char;
    pub fn ox_version() -> *const c_char;
    pub fn ox_model_from_xml_string(xml: *const c_char, out: *mut *mut ox_model) -> ox_status;
    pub fn ox_model_from_xml_path(path: *const c_char, out: *mut *mut ox_model) -> ox_status;
    pub fn ox_model_free(m: *mut ox_model);
    pub fn ox_model_get_tables(m: *const ox_model) -> *const c_void;
    pub fn ox_model_int_table(m: *const ox_model, name: *const c_char, ptr: *mut *const i32, count: *mut i32) -> ox_status;
    pub fn ox_model_real_table(m: *const ox_model, name: *const c_char, ptr: *mut *const c_double, count: *mut i32) -> ox_status;
    pub fn ox_model_size(m: *const ox_model, name: *const c_char) -> i32;
    pub fn ox_model_name2id(m: *const ox_model, objtype: i32, name: *const c_char) -> i32;
    pub fn ox_model_id2name(m: *const ox_model, objtype: i32, id: i32) -> *const c_char;
    pub fn ox_batch_config_default(cfg: *mut ox_batch_config);
    pub fn ox_batch_create(m: *const ox_model, cfg: *const ox_batch_config, out: *mut *mut ox_batch) -> ox_status;
    pub fn ox_batch_free(b: *mut ox_batch);
    pub fn ox_batch_nenv(b: *const ox_batch) -> i32;
    pub fn ox_batch_stream(b: *const ox_batch) -> *mut c_void;
    pub fn ox_batch_step(b: *mut ox_batch, nsteps: i32) -> ox_status;
    pub fn ox_batch_forward(b: *mut ox_batch) -> ox_status;
    pub fn ox_batch_step_io(b: *mut ox_batch, ctrl: *const c_void, qpos: *mut c_void, qvel: *mut c_void, dtype: i32, mem: i32) -> ox_status;
    pub fn ox_batch_reset(b: *mut ox_batch, host_mask_or_null: *const u8) -> ox_status;
    pub fn ox_batch_sync(b: *mut ox_batch) -> ox_status;
    pub fn ox_batch_ctrl_philox(b: *mut ox_batch, enable: i32, seed: u64) -> ox_status;
    pub fn ox_batch_set_step_counter(b: *mut ox_batch, step: i64) -> ox_status;
    pub fn ox_batch_field_size(b: *const ox_batch, field: i32) -> i32;
    pub fn ox_batch_get(b: *mut ox_batch, field: i32, buf: *mut c_void, dtype: i32, mem: i32, layout: i32) -> ox_status;
    pub fn ox_batch_get_many(b: *mut ox_batch, nfields: i32, fields: *const i32, bufs: *const *mut c_void, dtype: i32, mem: i32, layout: i32) -> ox_status;
    pub fn ox_batch_set(b: *mut ox_batch, field: i32, buf: *const c_void, dtype: i32, mem: i32, layout: i32) -> ox_status;
    pub fn ox_batch_get1(b: *mut ox_batch, field: i32, env: i32, offset: i32, count: i32, out: *mut c_double) -> ox_status;
    pub fn ox_batch_set1(b: *mut ox_batch, field: i32, env: i32, offset: i32, count: i32, inp: *const c_double) -> ox_status;
    pub fn ox_batch_get1_int(b: *mut ox_batch, field: i32, env: i32, offset: i32, count: i32, out: *mut i32) -> ox_status;
    pub fn ox_batch_stats(b: *mut ox_batch, out4: *mut c_double) -> ox_status;
    pub fn ox_batch_launch_count(b: *const ox_batch) -> i64;
    pub fn ox_batch_kernel_name(b: *const ox_batch) -> *const c_char;
    pub fn ox_spec_count() -> i32;
    pub fn ox_spec_name(i: i32) -> *const c_char;
    pub fn ox_batch_stage_times(b: *mut ox_batch, reps: i32, out_ms: *mut c_double, nstage: *mut i32) -> ox_status;
    pub fn ox_stage_name(i: i32) -> *const c_char;
    // N1: batched Environment / Task (src/lib.rs:8-88)
    pub fn ox_task_spec_default(spec: *mut ox_task_spec);
    pub fn ox_env_create(b: *mut ox_batch, spec: *const ox_task_spec, out: *mut *mut ox_env) -> ox_status;
    pub fn ox_env_free(e: *mut ox_env);
    pub fn ox_env_obs_dim(e: *const ox_env) -> i32;
    pub fn ox_env_reset(e: *mut ox_env, obs: *mut c_void, dtype: i32, mem: i32) -> ox_status;
    pub fn ox_env_step(e: *mut ox_env, action: *const c_void, obs: *mut c_void, reward: *mut c_void, discount: *mut c_void,
                       finished: *mut u8, dtype: i32, mem: i32) -> ox_status;
    pub fn ox_env_stats(e: *mut ox_env, out3: *mut c_double) -> ox_status;
}
#[allow(unused)] fn _types(_: c_int) {}
