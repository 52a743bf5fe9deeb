"""The Rust crates under rust/ cannot be compiled here (no cargo / rustc in the image), so what CAN be checked is checked:
the FFI crate is a mechanical transcription of include/ox_b200.h (regenerated and compared), and the safe wrapper declares
every method of the reference's `Physics` (/root/reference/src/physics.rs:11-171, names and receivers listed below) - the
same check test_abi.py makes for the ctypes binding and test_cpp_api.py makes (by compiling) for the C++ mirror."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (method, receiver) of `impl Physics` / `impl Actuators` in the reference; receiver: static | ref (&self) | mut (&mut self)
REFERENCE_SURFACE = [
    ("from_xml", "static"), ("from_xml_string", "static"), ("model", "ref"), ("data", "ref"), ("data_mut", "mut"), ("model_data", "ref"),
    ("model_datamut", "mut"), ("step", "mut"), ("forward", "mut"), ("reset", "mut"), ("object_id", "ref"), ("object_name", "ref"),
    ("set", "mut"), ("actuators", "mut"), ("time", "ref"), ("set_time", "mut"), ("ctrl", "ref"), ("set_ctrl", "mut"), ("act", "ref"),
    ("set_act", "mut"), ("qpos", "ref"), ("set_qpos", "mut"), ("qvel", "ref"), ("set_qvel", "mut"), ("qacc_warmstart", "ref"),
    ("set_qacc_warmstart", "mut"), ("plugin_state", "ref"), ("set_plugin_state", "mut"), ("qfrc_applied", "ref"), ("set_qfrc_applied", "mut"),
    ("xfrc_applied", "ref"), ("set_xfrc_applied", "mut"), ("eq_active", "ref"), ("set_eq_active", "mut"), ("mocap_pos", "ref"),
    ("set_mocap_pos", "mut"), ("mocap_quat", "ref"), ("set_mocap_quat", "mut"),
]


def test_sys_crate_is_the_generated_transcription_of_the_header():
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"]) == 0, \
        "rust/ox_b200-sys/src/lib.rs is stale: run python tools/gen_rust_sys.py"
    text = open(os.path.join(ROOT, "rust", "ox_b200-sys", "src", "lib.rs")).read()
    hdr = open(os.path.join(ROOT, "include", "ox_b200.h")).read()
    for name in set(re.findall(r"OX_API[^;(]*?\b(ox_[a-z0-9_]+)\s*\(", hdr)):
        assert f"pub fn {name}(" in text, name


def test_safe_wrapper_declares_the_reference_surface_with_the_reference_receivers():
    text = open(os.path.join(ROOT, "rust", "ox_b200", "src", "lib.rs")).read()
    physics = text[text.index("pub struct Physics {"):text.index("// ------------------------------------------------------------------------------------------------ Environment")]
    found = {}
    for name, recv in re.findall(r"pub fn (\w+)(?:<[^>]*>)?\(\s*(&mut self|&self|&'a mut self)?", physics):
        found[name] = {"&mut self": "mut", "&'a mut self": "mut", "&self": "ref", "": "static"}[recv]
    for name, recv in REFERENCE_SURFACE:
        assert name in found, f"rust wrapper lacks Physics::{name}"
        assert found[name] == recv, f"Physics::{name}: receiver {found[name]}, reference has {recv}"
    assert physics.count("{") == physics.count("}") and text.count("(") == text.count(")")   # at least balanced
    for extra in ("pub struct PhysicsGroup", "pub struct BatchedPhysics", "pub struct BatchedEnvironment", "pub trait Joint", "pub mod obj"):
        assert extra in text
