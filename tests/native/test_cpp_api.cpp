// Compiles the C++ host mirror (include/ox_b200.hpp) against libox_b200.so and exercises the paths that need no GPU;
// with a GPU it also steps a pendulum through the Physics API. Built and run by tests/test_cpp_api.py.
#include <cmath>
#include <cstdio>
#include <variant>

#include "../../include/ox_b200.hpp"

static const char* kPendulum = R"(<mujoco><compiler angle="radian"/><option timestep="0.002"/><worldbody><body name="pole">
<joint name="hinge" type="hinge" axis="0 1 0" damping="0.1"/><geom type="capsule" fromto="0 0 0 0 0 -0.5" size="0.02" contype="0" conaffinity="0"/>
</body></worldbody><actuator><motor name="torque" joint="hinge" ctrlrange="-1 1"/></actuator></mujoco>)";

int main() {
  using namespace ox_b200;
  Model m = Model::from_xml_string(kPendulum);
  if (m.tables().nq != 1 || m.tables().nu != 1) return 1;
  if (!m.object_id<OX_OBJ_JOINT>("hinge") || m.object_id<OX_OBJ_JOINT>("nope")) return 2;
  if (m.object_name(obj::Actuator{0}) != "torque") return 3;
  try { Model::from_xml_string("<mujoco><worldbody>"); return 4; } catch (const Error& e) { if (e.kind != Error::Kind::Mujoco) return 5; }
  try { Model::from_xml_string("<mujoco><option integrator=\"implicit\"/></mujoco>"); return 6; } catch (const Error& e) { if (e.kind != Error::Kind::Mjs) return 7; }
  try {
    Physics p = Physics::from_xml_string(kPendulum);
    auto hinge = *p.object_id<OX_OBJ_JOINT>("hinge");
    p.set_qpos<joint::Hinge>(hinge, {0.3});
    p.actuators().set(obj::Actuator{0}, 0.25);
    for (int i = 0; i < 100; i++) p.step();
    double q = p.qpos<joint::Hinge>(hinge)[0];
    if (!(std::fabs(p.time() - 0.2) < 1e-12) || !std::isfinite(q) || q == 0.3) return 8;
    try { p.qpos<joint::Slide>(hinge); return 9; } catch (const Error& e) { if (e.kind != Error::Kind::JointTypeNotMatch) return 10; }
    std::printf("gpu path ok: qpos after 100 steps = %.12f\n", q);
  } catch (const Error& e) {
    if (e.kind != Error::Kind::Cuda) return 11;   // no device: the product refuses (no CPU fallback)
    std::printf("no gpu: %s\n", e.what());
  }
  std::printf("cpp api ok\n");
  return 0;
}
