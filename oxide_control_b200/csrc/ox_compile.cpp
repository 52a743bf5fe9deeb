// MJCF-subset parser + model compiler (host, fp64). This is the setup side of the boundary:
// it stands in for mj_parseXMLString + mj_compile (+ mjs_getError) as called by
// Physics::from_xml_string, reference src/physics.rs:18-24, and mj_loadXML (src/physics.rs:13).
// Subset and defaults: SURVEY.md Appendix B. Anything outside the subset is a loud compile error,
// never silently ignored (cosmetic attributes such as rgba/material/group are the exception).
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <set>
#include <sstream>
#include <stdexcept>

#include "ox_hostmath.h"
#include "ox_model.h"
#include "ox_xml.h"

namespace ox {
namespace {

using AttrMap = std::map<std::string, std::string>;
constexpr double kPi = 3.14159265358979323846;

[[noreturn]] void cfail(const std::string& m) { throw CompileError(m); }
[[noreturn]] void pfail(const XmlElem& e, const std::string& m) {
  throw XmlError("XML error at line " + std::to_string(e.line) + " in <" + e.name + ">: " + m);
}

std::vector<double> parse_nums(const std::string& s, const std::string& ctx) {
  std::vector<double> v;
  const char* p = s.c_str();
  for (;;) {
    while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == ',') ++p;
    if (!*p) break;
    char* end = nullptr;
    double d = std::strtod(p, &end);
    if (end == p) throw XmlError("XML error: bad number in attribute " + ctx + "=\"" + s + "\"");
    v.push_back(d);
    p = end;
  }
  return v;
}

// ---- schema: known attributes per element (unknown => parse error, like MuJoCo's schema check) ----
const std::map<std::string, std::set<std::string>>& schema() {
  static const std::map<std::string, std::set<std::string>> s = {
      {"mujoco", {"model"}},
      {"compiler", {"angle", "coordinate", "inertiafromgeom", "autolimits", "eulerseq", "boundmass", "boundinertia",
                    "settotalmass", "balanceinertia", "strippath", "meshdir", "texturedir", "assetdir", "discardvisual",
                    "fusestatic", "inertiagrouprange", "usethread", "alignfree"}},
      {"option", {"timestep", "gravity", "integrator", "solver", "iterations", "tolerance", "ls_iterations", "ls_tolerance",
                  "cone", "impratio", "jacobian", "noslip_iterations", "noslip_tolerance", "ccd_iterations", "ccd_tolerance",
                  "wind", "magnetic", "density", "viscosity", "o_margin", "o_solref", "o_solimp", "o_friction", "apirate",
                  "sdf_iterations", "sdf_initpoints", "actuatorgroupdisable"}},
      {"flag", {"constraint", "equality", "frictionloss", "limit", "contact", "passive", "gravity", "clampctrl", "warmstart",
                "filterparent", "actuation", "refsafe", "sensor", "midphase", "eulerdamp", "autoreset", "nativeccd", "override",
                "energy", "fwdinv", "invdiscrete", "multiccd", "island", "spring", "damper"}},
      {"body", {"name", "pos", "quat", "euler", "axisangle", "xyaxes", "zaxis", "childclass", "mocap", "gravcomp", "user"}},
      {"inertial", {"pos", "quat", "euler", "axisangle", "xyaxes", "zaxis", "mass", "diaginertia", "fullinertia"}},
      {"joint", {"name", "class", "type", "pos", "axis", "range", "limited", "damping", "stiffness", "springref", "armature",
                 "ref", "margin", "solreflimit", "solimplimit", "frictionloss", "solreffriction", "solimpfriction", "group",
                 "springdamper", "actuatorfrcrange", "actuatorfrclimited", "actuatorgravcomp", "user"}},
      {"freejoint", {"name", "group", "align"}},
      {"geom", {"name", "class", "type", "size", "fromto", "pos", "quat", "euler", "axisangle", "xyaxes", "zaxis", "density",
                "mass", "friction", "condim", "contype", "conaffinity", "margin", "gap", "solref", "solimp", "solmix",
                "priority", "rgba", "material", "group", "shellinertia", "fluidshape", "fluidcoef", "user", "mesh", "hfield",
                "fitscale"}},
      {"site", {"name", "class", "type", "size", "fromto", "pos", "quat", "euler", "axisangle", "xyaxes", "zaxis", "rgba",
                "material", "group", "user"}},
      {"actuator_common", {"name", "class", "joint", "gear", "ctrlrange", "ctrllimited", "forcerange", "forcelimited", "group",
                           "kp", "kv", "gaintype", "biastype", "gainprm", "biasprm", "dyntype", "dynprm", "actrange",
                           "actlimited", "user", "tendon", "site", "body", "jointinparent", "lengthrange", "cranklength",
                           "slidersite", "cranksite", "refsite", "actdim", "actearly", "inheritrange", "dampratio", "timeconst"}},
      {"exclude", {"name", "body1", "body2"}},
      {"tendon_fixed", {"name", "class", "group", "limited", "range", "solreflimit", "solimplimit", "solreffriction", "solimpfriction", "margin",
                        "frictionloss", "springlength", "stiffness", "damping", "user", "width", "material", "rgba"}},
      {"tendon_joint", {"joint", "coef"}},
      {"tendon_site", {"site"}},
      {"equality_common", {"name", "class", "active", "solref", "solimp", "body1", "body2", "anchor", "joint1", "joint2", "polycoef", "relpose", "torquescale"}},
      {"sensor_common", {"name", "joint", "actuator", "site", "body", "tendon", "objtype", "objname", "reftype", "refname", "cutoff",
                         "noise", "user"}},
  };
  return s;
}

void check_attrs(const XmlElem& e, const std::string& schema_key) {
  auto it = schema().find(schema_key);
  if (it == schema().end()) return;
  for (auto& a : e.attrs)
    if (!it->second.count(a.first)) pfail(e, "unrecognized attribute '" + a.first + "'");
}

struct Ctx {
  bool degree = true;
  bool autolimits = true;
  int inertiafromgeom = 2;  // 0 false, 1 true, 2 auto
  std::string eulerseq = "xyz";
  double boundmass = 0, boundinertia = 0;
  double settotalmass = -1;   // > 0: rescale every body's mass and inertia so that the model weighs this much
  // defaults: class -> element tag -> attrs
  std::map<std::string, std::map<std::string, AttrMap>> defaults;
};

struct Attrs {  // merged attribute view of an element
  AttrMap m;
  const XmlElem* e;
  bool has(const std::string& k) const { return m.count(k) != 0; }
  const std::string& str(const std::string& k) const { return m.at(k); }
  std::string str_or(const std::string& k, const std::string& d) const { return has(k) ? m.at(k) : d; }
  std::vector<double> nums(const std::string& k) const { return parse_nums(m.at(k), k); }
  double num(const std::string& k, double d) const {
    if (!has(k)) return d;
    auto v = nums(k);
    if (v.size() != 1) pfail(*e, "attribute '" + k + "' expects one number");
    return v[0];
  }
  void vec(const std::string& k, double* out, int n, bool pad_ok = false) const {
    if (!has(k)) return;
    auto v = nums(k);
    if ((int)v.size() > n || (!pad_ok && (int)v.size() != n)) pfail(*e, "attribute '" + k + "' expects " + std::to_string(n) + " numbers");
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
  }
  // tri-state bool: -1 absent/auto
  int boolean(const std::string& k) const {
    if (!has(k)) return -1;
    const std::string& s = m.at(k);
    if (s == "true") return 1;
    if (s == "false") return 0;
    if (s == "auto") return -1;
    pfail(*e, "attribute '" + k + "' must be true/false/auto");
  }
};

Attrs merged(const Ctx& c, const XmlElem& e, const std::string& tag, const std::string& childclass) {
  Attrs a;
  a.e = &e;
  std::string cls = "main";
  if (const std::string* s = e.attr("class")) cls = *s;
  else if (!childclass.empty()) cls = childclass;
  auto ci = c.defaults.find(cls);
  if (ci == c.defaults.end()) {
    if (cls != "main") pfail(e, "unknown default class '" + cls + "'");
  } else {
    auto ti = ci->second.find(tag);
    if (ti != ci->second.end()) a.m = ti->second;
  }
  for (auto& kv : e.attrs) a.m[kv.first] = kv.second;
  return a;
}

bool is_actuator_tag(const std::string& n) {
  return n == "motor" || n == "position" || n == "velocity" || n == "general";
}

void parse_defaults(Ctx& c, const XmlElem& d, const std::string& parent_cls, bool top) {
  std::string cls = "main";
  if (const std::string* s = d.attr("class")) cls = *s;
  else if (!top) pfail(d, "nested <default> requires a class");
  if (!top && c.defaults.count(cls)) pfail(d, "repeated default class '" + cls + "'");
  if (!parent_cls.empty() && cls != parent_cls) c.defaults[cls] = c.defaults[parent_cls];  // inherit
  auto& dst = c.defaults[cls];
  for (auto& ch : d.children) {
    if (ch->name == "default") continue;
    std::string tag = ch->name;
    if (tag == "geom") check_attrs(*ch, "geom");
    else if (tag == "joint") check_attrs(*ch, "joint");
    else if (tag == "site") check_attrs(*ch, "site");
    else if (tag == "tendon") check_attrs(*ch, "tendon_fixed");
    else if (is_actuator_tag(tag)) { check_attrs(*ch, "actuator_common"); tag = "actuator"; }
    else if (tag == "camera" || tag == "light" || tag == "material" || tag == "mesh") continue;  // cosmetic
    else cfail("unsupported element <" + ch->name + "> inside <default> (line " + std::to_string(ch->line) + ")");
    for (auto& kv : ch->attrs) dst[tag][kv.first] = kv.second;
  }
  for (auto& ch : d.children)
    if (ch->name == "default") parse_defaults(c, *ch, cls, false);
}

// orientation attributes -> quaternion
void orientation(const Ctx& c, const Attrs& a, double* quat) {
  int n = a.has("quat") + a.has("euler") + a.has("axisangle") + a.has("xyaxes") + a.has("zaxis");
  if (n > 1) pfail(*a.e, "multiple orientation specifiers");
  const double ang = c.degree ? kPi / 180.0 : 1.0;
  if (a.has("quat")) {
    a.vec("quat", quat, 4);
    hm::normalize4(quat);
  } else if (a.has("euler")) {
    double e[3] = {0, 0, 0};
    a.vec("euler", e, 3);
    double q[4] = {1, 0, 0, 0};
    for (int i = 0; i < 3; i++) {
      char ch = c.eulerseq[i];
      double ax[3] = {0, 0, 0};
      char lo = (char)std::tolower(ch);
      ax[lo - 'x'] = 1;
      double r[4], t[4];
      hm::axisangle2quat(r, ax, e[i] * ang);
      if (ch == lo) hm::mulquat(t, q, r);  // intrinsic (moving axes)
      else hm::mulquat(t, r, q);           // extrinsic
      std::memcpy(q, t, sizeof q);
    }
    hm::normalize4(q);
    std::memcpy(quat, q, sizeof q);
  } else if (a.has("axisangle")) {
    double v[4] = {0, 0, 1, 0};
    a.vec("axisangle", v, 4);
    hm::normalize3(v);
    hm::axisangle2quat(quat, v, v[3] * ang);
  } else if (a.has("zaxis")) {
    double v[3] = {0, 0, 1};
    a.vec("zaxis", v, 3);
    hm::z2quat(quat, v);
  } else if (a.has("xyaxes")) {
    double v[6] = {1, 0, 0, 0, 1, 0};
    a.vec("xyaxes", v, 6);
    double x[3] = {v[0], v[1], v[2]}, y[3] = {v[3], v[4], v[5]}, z[3];
    hm::normalize3(x);
    double d = hm::dot3(x, y);
    for (int i = 0; i < 3; i++) y[i] -= d * x[i];
    hm::normalize3(y);
    hm::cross(z, x, y);
    double R[9] = {x[0], y[0], z[0], x[1], y[1], z[1], x[2], y[2], z[2]};
    hm::mat2quat(quat, R);
  }
}

struct GeomDef {
  std::string name;
  int type = OX_GEOM_SPHERE;
  double size[3] = {0, 0, 0}, pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0};
  double density = 1000, mass = -1;
  double friction[3] = {1, 0.005, 0.0001};
  int condim = 3, contype = 1, conaffinity = 1, priority = 0;
  double margin = 0, gap = 0, solmix = 1;
  double solref[2] = {0.02, 1}, solimp[5] = {0.9, 0.95, 0.001, 0.5, 2};
  int body = 0;
};
struct SiteDef {
  std::string name;
  double pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0};
  int body = 0;
  int type = OX_GEOM_SPHERE;                       // site shapes matter to the touch sensor only (its sensing volume)
  double size[3] = {0.005, 0.005, 0.005};
};
struct JointDef {
  std::string name;
  int type = OX_JNT_HINGE;
  double pos[3] = {0, 0, 0}, axis[3] = {0, 0, 1};
  double range[2] = {0, 0};
  int limited = 0;
  double damping = 0, stiffness = 0, springref = 0, armature = 0, ref = 0, margin = 0;
  double solref[2] = {0.02, 1}, solimp[5] = {0.9, 0.95, 0.001, 0.5, 2};
  double frictionloss = 0, solref_fri[2] = {0.02, 1}, solimp_fri[5] = {0.9, 0.95, 0.001, 0.5, 2};
  int body = 0;
};
struct BodyDef {
  std::string name;
  int parent = 0;
  double pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0};
  bool has_inertial = false, mocap = false;
  double ipos[3] = {0, 0, 0}, iquat[4] = {1, 0, 0, 0}, mass = 0, inertia[3] = {0, 0, 0};
  double gravcomp = 0;   // fraction of the body's weight cancelled by a passive force at its com
  std::vector<int> joints, geoms;
};

struct Builder {
  Ctx c;
  std::vector<BodyDef> bodies;
  std::vector<JointDef> joints;
  std::vector<GeomDef> geoms;
  std::vector<SiteDef> sites;
  std::set<std::pair<int, int>> excludes;
};

int geom_type_from(const std::string& s, const XmlElem& e) {
  static const std::map<std::string, int> m = {{"plane", OX_GEOM_PLANE}, {"hfield", OX_GEOM_HFIELD}, {"sphere", OX_GEOM_SPHERE},
                                               {"capsule", OX_GEOM_CAPSULE}, {"ellipsoid", OX_GEOM_ELLIPSOID},
                                               {"cylinder", OX_GEOM_CYLINDER}, {"box", OX_GEOM_BOX}, {"mesh", OX_GEOM_MESH}};
  auto it = m.find(s);
  if (it == m.end()) pfail(e, "unknown geom type '" + s + "'");
  return it->second;
}

void clamp_solimp(double* s) {
  s[0] = std::min(OX_MAXIMP, std::max(OX_MINIMP, s[0]));
  s[1] = std::min(OX_MAXIMP, std::max(OX_MINIMP, s[1]));
  s[2] = std::max(0.0, s[2]);
  s[3] = std::min(OX_MAXIMP, std::max(OX_MINIMP, s[3]));
  s[4] = std::max(1.0, s[4]);
}

void parse_geom(Builder& B, const XmlElem& e, int body, const std::string& childclass) {
  check_attrs(e, "geom");
  Attrs a = merged(B.c, e, "geom", childclass);
  GeomDef g;
  g.body = body;
  g.name = a.str_or("name", "");
  g.type = geom_type_from(a.str_or("type", "sphere"), e);
  if (g.type == OX_GEOM_MESH || g.type == OX_GEOM_HFIELD || a.has("mesh") || a.has("hfield"))
    cfail("geom '" + g.name + "': mesh/hfield geoms are outside the supported MJCF subset");
  a.vec("size", g.size, 3, true);
  a.vec("pos", g.pos, 3);
  orientation(B.c, a, g.quat);
  if (a.has("fromto")) {
    if (g.type != OX_GEOM_CAPSULE && g.type != OX_GEOM_CYLINDER && g.type != OX_GEOM_BOX && g.type != OX_GEOM_ELLIPSOID)
      pfail(e, "fromto requires capsule, cylinder, box or ellipsoid");
    double ft[6];
    a.vec("fromto", ft, 6);
    double v[3] = {ft[0] - ft[3], ft[1] - ft[4], ft[2] - ft[5]};
    double len = hm::norm3(v);
    if (len < 1e-15) cfail("geom '" + g.name + "': fromto points too close");
    for (int i = 0; i < 3; i++) g.pos[i] = 0.5 * (ft[i] + ft[i + 3]);
    hm::z2quat(g.quat, v);
    if (g.type == OX_GEOM_CAPSULE || g.type == OX_GEOM_CYLINDER) g.size[1] = len / 2;
    else g.size[2] = len / 2;
  }
  g.density = a.num("density", g.density);
  g.mass = a.num("mass", -1);
  if (a.has("shellinertia") && a.str("shellinertia") == "true") cfail("geom '" + g.name + "': shellinertia is outside the supported subset");
  if (a.has("fluidshape") && a.str("fluidshape") != "none")
    cfail("geom '" + g.name + "': fluidshape='" + a.str("fluidshape") + "' (ellipsoid fluid model) is outside the supported subset (inertia-box model)");
  a.vec("friction", g.friction, 3, true);
  g.condim = (int)a.num("condim", g.condim);
  g.contype = (int)a.num("contype", g.contype);
  g.conaffinity = (int)a.num("conaffinity", g.conaffinity);
  g.priority = (int)a.num("priority", g.priority);
  g.margin = a.num("margin", 0);
  g.gap = a.num("gap", 0);
  g.solmix = a.num("solmix", 1);
  a.vec("solref", g.solref, 2, true);
  a.vec("solimp", g.solimp, 5, true);
  clamp_solimp(g.solimp);
  if (g.condim != 1 && g.condim != 3 && g.condim != 4 && g.condim != 6) pfail(e, "condim must be 1, 3, 4 or 6");
  // size checks
  auto need = [&](int n) {
    for (int i = 0; i < n; i++)
      if (!(g.size[i] > 0)) cfail("geom '" + g.name + "': size " + std::to_string(i) + " must be positive");
  };
  switch (g.type) {
    case OX_GEOM_SPHERE: need(1); break;
    case OX_GEOM_CAPSULE: case OX_GEOM_CYLINDER: need(2); break;
    case OX_GEOM_BOX: case OX_GEOM_ELLIPSOID: need(3); break;
    default: break;
  }
  B.bodies[body].geoms.push_back((int)B.geoms.size());
  B.geoms.push_back(g);
}

void parse_site(Builder& B, const XmlElem& e, int body, const std::string& childclass) {
  check_attrs(e, "site");
  Attrs a = merged(B.c, e, "site", childclass);
  SiteDef s;
  s.body = body;
  s.name = a.str_or("name", "");
  a.vec("pos", s.pos, 3);
  orientation(B.c, a, s.quat);
  s.type = geom_type_from(a.str_or("type", "sphere"), e);
  a.vec("size", s.size, 3, true);
  if (a.has("fromto")) {
    double ft[6];
    a.vec("fromto", ft, 6);
    double v[3] = {ft[0] - ft[3], ft[1] - ft[4], ft[2] - ft[5]};
    for (int i = 0; i < 3; i++) s.pos[i] = 0.5 * (ft[i] + ft[i + 3]);
    hm::z2quat(s.quat, v);
    const double len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (s.type == OX_GEOM_CAPSULE || s.type == OX_GEOM_CYLINDER) s.size[1] = len / 2; else s.size[2] = len / 2;
  }
  B.sites.push_back(s);
}

void parse_joint(Builder& B, const XmlElem& e, int body, const std::string& childclass, bool freejoint) {
  JointDef j;
  j.body = body;
  const double ang = B.c.degree ? kPi / 180.0 : 1.0;
  if (freejoint) {
    check_attrs(e, "freejoint");
    if (const std::string* n = e.attr("name")) j.name = *n;
    j.type = OX_JNT_FREE;
  } else {
    check_attrs(e, "joint");
    Attrs a = merged(B.c, e, "joint", childclass);
    j.name = a.str_or("name", "");
    std::string t = a.str_or("type", "hinge");
    if (t == "free") j.type = OX_JNT_FREE;
    else if (t == "ball") j.type = OX_JNT_BALL;
    else if (t == "slide") j.type = OX_JNT_SLIDE;
    else if (t == "hinge") j.type = OX_JNT_HINGE;
    else pfail(e, "unknown joint type '" + t + "'");
    a.vec("pos", j.pos, 3);
    a.vec("axis", j.axis, 3);
    bool rot = (j.type == OX_JNT_HINGE || j.type == OX_JNT_BALL);
    bool has_range = a.has("range");
    if (has_range) {
      a.vec("range", j.range, 2);
      if (rot) { j.range[0] *= ang; j.range[1] *= ang; }
    }
    int lim = a.boolean("limited");
    if (lim < 0) {
      if (has_range && !B.c.autolimits) cfail("joint '" + j.name + "': range given but limited unspecified and autolimits=false");
      lim = has_range ? 1 : 0;
    }
    j.limited = lim;
    if (j.limited && !(j.range[0] < j.range[1])) cfail("joint '" + j.name + "': range[0] must be < range[1]");
    j.damping = a.num("damping", 0);
    j.stiffness = a.num("stiffness", 0);
    j.armature = a.num("armature", 0);
    j.margin = a.num("margin", 0);
    j.ref = a.num("ref", 0) * (j.type == OX_JNT_HINGE ? ang : 1.0);
    j.springref = a.num("springref", 0) * (j.type == OX_JNT_HINGE ? ang : 1.0);
    a.vec("solreflimit", j.solref, 2, true);
    a.vec("solimplimit", j.solimp, 5, true);
    clamp_solimp(j.solimp);
    j.frictionloss = a.num("frictionloss", 0);
    if (j.frictionloss < 0) cfail("joint '" + j.name + "': frictionloss must be >= 0");
    a.vec("solreffriction", j.solref_fri, 2, true);
    a.vec("solimpfriction", j.solimp_fri, 5, true);
    clamp_solimp(j.solimp_fri);
    if (a.has("actuatorfrcrange")) cfail("joint '" + j.name + "': actuatorfrcrange is outside the supported subset");
    if (a.has("actuatorfrclimited") && a.str("actuatorfrclimited") == "true") cfail("joint '" + j.name + "': actuatorfrclimited is outside the supported subset");
    if (a.has("actuatorgravcomp") && a.str("actuatorgravcomp") == "true") cfail("joint '" + j.name + "': actuatorgravcomp is outside the supported subset");
    if (a.has("springdamper")) {
      double sd[2] = {0, 0};
      a.vec("springdamper", sd, 2);
      if (sd[0] != 0 || sd[1] != 0) cfail("joint '" + j.name + "': springdamper is outside the supported subset (give stiffness and damping)");
    }
    if (j.type == OX_JNT_FREE || j.type == OX_JNT_BALL) { j.axis[0] = j.axis[1] = 0; j.axis[2] = 1; }
    else if (hm::normalize3(j.axis) < 1e-15) cfail("joint '" + j.name + "': zero axis");
  }
  B.bodies[body].joints.push_back((int)B.joints.size());
  B.joints.push_back(j);
}

void parse_body(Builder& B, const XmlElem& e, int parent, std::string childclass, bool world) {
  int id = parent;
  if (!world) {
    check_attrs(e, "body");
    Attrs a;
    a.e = &e;
    for (auto& kv : e.attrs) a.m[kv.first] = kv.second;
    BodyDef b;
    b.parent = parent;
    b.name = a.str_or("name", "");
    a.vec("pos", b.pos, 3);
    orientation(B.c, a, b.quat);
    b.gravcomp = a.num("gravcomp", 0);
    if (a.has("mocap") && a.str("mocap") == "true") {
      if (parent != 0) cfail("body '" + b.name + "': a mocap body must be a child of the world");
      b.mocap = true;
    }
    if (a.has("childclass")) {
      childclass = a.str("childclass");
      if (!B.c.defaults.count(childclass)) pfail(e, "unknown default class '" + childclass + "'");
    }
    id = (int)B.bodies.size();
    B.bodies.push_back(b);
  }
  // MuJoCo order: the body's own elements first, child bodies afterwards (depth-first)
  for (auto& ch : e.children) {
    const std::string& n = ch->name;
    if (n == "body") continue;
    if (n == "inertial") {
      if (world) pfail(*ch, "inertial not allowed in worldbody");
      check_attrs(*ch, "inertial");
      Attrs a;
      a.e = ch.get();
      for (auto& kv : ch->attrs) a.m[kv.first] = kv.second;
      BodyDef& b = B.bodies[id];
      b.has_inertial = true;
      a.vec("pos", b.ipos, 3);
      orientation(B.c, a, b.iquat);
      if (!a.has("mass")) pfail(*ch, "inertial requires mass");
      b.mass = a.num("mass", 0);
      if (a.has("diaginertia")) a.vec("diaginertia", b.inertia, 3);
      else if (a.has("fullinertia")) {
        double f[6];
        a.vec("fullinertia", f, 6);  // xx yy zz xy xz yz
        double A[9] = {f[0], f[3], f[4], f[3], f[1], f[5], f[4], f[5], f[2]}, ev[3], R[9];
        hm::eig3(ev, R, A);
        std::memcpy(b.inertia, ev, sizeof ev);
        hm::mat2quat(b.iquat, R);
      } else pfail(*ch, "inertial requires diaginertia or fullinertia");
    } else if (n == "joint") {
      if (world) pfail(*ch, "joint not allowed in worldbody");
      parse_joint(B, *ch, id, childclass, false);
    } else if (n == "freejoint") {
      if (world) pfail(*ch, "freejoint not allowed in worldbody");
      parse_joint(B, *ch, id, childclass, true);
    } else if (n == "geom") parse_geom(B, *ch, id, childclass);
    else if (n == "site") parse_site(B, *ch, id, childclass);
    else if (n == "camera" || n == "light") continue;
    else cfail("unsupported element <" + n + "> in body (line " + std::to_string(ch->line) + ")");
  }
  for (auto& ch : e.children)
    if (ch->name == "body") parse_body(B, *ch, id, childclass, false);
}

// ---- mass properties ----
double geom_volume(const GeomDef& g) {
  const double* s = g.size;
  switch (g.type) {
    case OX_GEOM_SPHERE: return 4.0 / 3.0 * kPi * s[0] * s[0] * s[0];
    case OX_GEOM_CAPSULE: { double h = 2 * s[1]; return kPi * (s[0] * s[0] * h + 4.0 / 3.0 * s[0] * s[0] * s[0]); }
    case OX_GEOM_CYLINDER: return kPi * s[0] * s[0] * 2 * s[1];
    case OX_GEOM_ELLIPSOID: return 4.0 / 3.0 * kPi * s[0] * s[1] * s[2];
    case OX_GEOM_BOX: return 8 * s[0] * s[1] * s[2];
    default: return 0;
  }
}
void geom_inertia(const GeomDef& g, double mass, double* I) {
  const double* s = g.size;
  I[0] = I[1] = I[2] = 0;
  switch (g.type) {
    case OX_GEOM_SPHERE: I[0] = I[1] = I[2] = 2 * mass * s[0] * s[0] / 5; break;
    case OX_GEOM_CAPSULE: {
      double r = s[0], h = 2 * s[1];
      double sm = mass * (4.0 / 3.0 * r * r * r) / (r * r * h + 4.0 / 3.0 * r * r * r), cm = mass - sm;
      I[0] = I[1] = cm * (3 * r * r + h * h) / 12;
      I[2] = cm * r * r / 2;
      double si = 2 * sm * r * r / 5;
      I[0] += si + sm * h * (3 * r + 2 * h) / 8;
      I[1] = I[0];
      I[2] += si;
      break;
    }
    case OX_GEOM_CYLINDER: {
      double r = s[0], h = 2 * s[1];
      I[0] = I[1] = mass * (3 * r * r + h * h) / 12;
      I[2] = mass * r * r / 2;
      break;
    }
    case OX_GEOM_ELLIPSOID:
      I[0] = mass * (s[1] * s[1] + s[2] * s[2]) / 5;
      I[1] = mass * (s[0] * s[0] + s[2] * s[2]) / 5;
      I[2] = mass * (s[0] * s[0] + s[1] * s[1]) / 5;
      break;
    case OX_GEOM_BOX:
      I[0] = mass * (s[1] * s[1] + s[2] * s[2]) / 3;
      I[1] = mass * (s[0] * s[0] + s[2] * s[2]) / 3;
      I[2] = mass * (s[0] * s[0] + s[1] * s[1]) / 3;
      break;
    default: break;
  }
}

void body_inertia_from_geoms(Builder& B, BodyDef& b) {
  double mtot = 0, com[3] = {0, 0, 0};
  std::vector<double> gm;
  for (int gi : b.geoms) {
    const GeomDef& g = B.geoms[gi];
    double m = g.mass >= 0 ? g.mass : g.density * geom_volume(g);
    if (g.type == OX_GEOM_PLANE) m = 0;
    gm.push_back(m);
    mtot += m;
    for (int k = 0; k < 3; k++) com[k] += m * g.pos[k];
  }
  if (mtot < 1e-15) { b.mass = 0; return; }
  for (int k = 0; k < 3; k++) com[k] /= mtot;
  double A[9] = {0};
  for (size_t n = 0; n < b.geoms.size(); n++) {
    const GeomDef& g = B.geoms[b.geoms[n]];
    double m = gm[n];
    if (m == 0) continue;
    double I[3], R[9];
    geom_inertia(g, m, I);
    hm::quat2mat(R, g.quat);
    double d[3] = {g.pos[0] - com[0], g.pos[1] - com[1], g.pos[2] - com[2]};
    double d2 = hm::dot3(d, d);
    for (int r = 0; r < 3; r++)
      for (int cc = 0; cc < 3; cc++) {
        double v = 0;
        for (int k = 0; k < 3; k++) v += R[3 * r + k] * I[k] * R[3 * cc + k];
        v += m * ((r == cc ? d2 : 0) - d[r] * d[cc]);
        A[3 * r + cc] += v;
      }
  }
  double ev[3], R[9];
  hm::eig3(ev, R, A);
  b.mass = mtot;
  std::memcpy(b.ipos, com, sizeof com);
  std::memcpy(b.inertia, ev, sizeof ev);
  hm::mat2quat(b.iquat, R);
}

struct ActDef {
  std::string name;
  int joint = -1, tendon = -1;
  double gear = 1;
  int gaintype = OX_GAIN_FIXED, biastype = OX_BIAS_NONE;
  double gainprm[3] = {1, 0, 0}, biasprm[3] = {0, 0, 0};
  int ctrllimited = 0, forcelimited = 0;
  double ctrlrange[2] = {0, 0}, forcerange[2] = {0, 0};
  int dyntype = OX_DYN_NONE, actlimited = 0;
  double dynprm[3] = {1, 0, 0}, actrange[2] = {0, 0};
};
struct SensorDef {
  std::string name;
  int type, objtype, objid, dim;
};

int find_name(const std::vector<std::string>& v, const std::string& n) {
  if (n.empty()) return -1;
  for (size_t i = 0; i < v.size(); i++)
    if (v[i] == n) return (int)i;
  return -1;
}

}  // namespace

ox_model* compile_mjcf(const std::string& xml) {
  XmlParser parser(xml);
  std::unique_ptr<XmlElem> root = parser.parse();
  if (root->name != "mujoco") pfail(*root, "root element must be <mujoco>");
  check_attrs(*root, "mujoco");

  Builder B;
  B.bodies.emplace_back();
  B.bodies[0].name = "world";
  auto M = std::make_unique<ox_model>();
  ox_model_tables& t = M->t;
  if (const std::string* n = root->attr("model")) M->model_name = *n;

  // option defaults
  t.timestep = 0.002;
  t.gravity[0] = 0; t.gravity[1] = 0; t.gravity[2] = -9.81;
  t.integrator = OX_INT_EULER;
  t.solver = OX_SOL_NEWTON;
  t.cone = 0;
  t.iterations = 100;
  t.tolerance = 1e-8;
  t.ls_iterations = 50;
  t.ls_tolerance = 0.01;
  t.impratio = 1;
  t.disableflags = 0;
  t.noslip_iterations = 0;
  t.noslip_tolerance = 1e-6;
  t.density = 0; t.viscosity = 0; t.wind[0] = t.wind[1] = t.wind[2] = 0;

  // pass 1: compiler, option, defaults (must precede use regardless of document order)
  for (auto& ch : root->children) {
    if (ch->name == "compiler") {
      check_attrs(*ch, "compiler");
      if (auto* s = ch->attr("angle")) {
        if (*s == "radian") B.c.degree = false;
        else if (*s == "degree") B.c.degree = true;
        else pfail(*ch, "angle must be radian or degree");
      }
      if (auto* s = ch->attr("coordinate"))
        if (*s != "local") cfail("compiler coordinate='" + *s + "' is not supported (local only)");
      if (auto* s = ch->attr("inertiafromgeom")) B.c.inertiafromgeom = (*s == "true") ? 1 : (*s == "false") ? 0 : 2;
      if (auto* s = ch->attr("autolimits")) B.c.autolimits = (*s != "false");
      if (auto* s = ch->attr("eulerseq")) {
        if (s->size() != 3) pfail(*ch, "eulerseq must have 3 characters");
        for (char cc : *s)
          if (std::string("xyzXYZ").find(cc) == std::string::npos) pfail(*ch, "eulerseq must use xyzXYZ");
        B.c.eulerseq = *s;
      }
      if (auto* s = ch->attr("settotalmass")) B.c.settotalmass = std::stod(*s);
      if (ch->attr("inertiagrouprange")) cfail("compiler inertiagrouprange is outside the supported subset");
      if (auto* s = ch->attr("balanceinertia")) if (*s == "true") cfail("compiler balanceinertia is outside the supported subset");
      if (auto* s = ch->attr("boundmass")) B.c.boundmass = std::stod(*s);
      if (auto* s = ch->attr("boundinertia")) B.c.boundinertia = std::stod(*s);
      if (B.c.boundmass < 0 || B.c.boundinertia < 0) cfail("compiler boundmass / boundinertia must be >= 0");
    }
  }
  for (auto& ch : root->children)
    if (ch->name == "default") parse_defaults(B.c, *ch, "", true);

  for (auto& ch : root->children) {
    const std::string& n = ch->name;
    if (n == "option") {
      check_attrs(*ch, "option");
      Attrs a;
      a.e = ch.get();
      for (auto& kv : ch->attrs) a.m[kv.first] = kv.second;
      t.timestep = a.num("timestep", t.timestep);
      a.vec("gravity", t.gravity, 3);
      t.iterations = (int)a.num("iterations", t.iterations);
      t.tolerance = a.num("tolerance", t.tolerance);
      t.ls_iterations = (int)a.num("ls_iterations", t.ls_iterations);
      t.ls_tolerance = a.num("ls_tolerance", t.ls_tolerance);
      t.impratio = a.num("impratio", t.impratio);
      t.noslip_iterations = (int)a.num("noslip_iterations", t.noslip_iterations);
      t.noslip_tolerance = a.num("noslip_tolerance", t.noslip_tolerance);
      if (t.noslip_iterations < 0) cfail("noslip_iterations must be >= 0");
      t.density = a.num("density", t.density);
      t.viscosity = a.num("viscosity", t.viscosity);
      a.vec("wind", t.wind, 3);
      if (t.density < 0 || t.viscosity < 0) cfail("option density / viscosity must be >= 0");
      if (a.has("actuatorgroupdisable") && a.str("actuatorgroupdisable").find_first_not_of(" \t\n") != std::string::npos)
        cfail("option actuatorgroupdisable is outside the supported subset");
      if (a.has("integrator")) {
        const std::string& s = a.str("integrator");
        if (s == "Euler") t.integrator = OX_INT_EULER;
        else if (s == "RK4") t.integrator = OX_INT_RK4;
        else if (s == "implicitfast") t.integrator = OX_INT_IMPLICITFAST;
        else if (s == "implicit") cfail("integrator 'implicit' (full RNE velocity derivatives, non-symmetric solve) is outside the supported subset (Euler, RK4, implicitfast)");
        else pfail(*ch, "unknown integrator '" + s + "'");
      }
      if (a.has("solver")) {
        const std::string& s = a.str("solver");
        if (s == "Newton") t.solver = OX_SOL_NEWTON;
        else if (s == "CG") t.solver = OX_SOL_CG;
        else if (s == "PGS") t.solver = OX_SOL_PGS;
        else pfail(*ch, "unknown solver '" + s + "'");
      }
      if (a.has("cone")) {
        const std::string& s = a.str("cone");
        if (s == "elliptic") t.cone = OX_CONE_ELLIPTIC;
        else if (s != "pyramidal") pfail(*ch, "unknown cone '" + s + "'");
      }
      if (t.cone == OX_CONE_ELLIPTIC && (t.solver == OX_SOL_PGS || t.noslip_iterations > 0))
        cfail("cone elliptic with the PGS solver or the noslip pass is outside the supported subset (Newton, CG)");
      if ((t.density > 0 || t.viscosity > 0) && t.integrator == OX_INT_IMPLICITFAST)
        cfail("fluid forces (option density / viscosity) with the implicitfast integrator (their velocity derivative is not built) are outside the supported subset");
      if (!(t.timestep > 0)) cfail("timestep must be positive");
      if (t.impratio <= 0) cfail("impratio must be positive");
      for (auto& f : ch->children) {
        if (f->name != "flag") pfail(*f, "unexpected element inside <option>");
        check_attrs(*f, "flag");
        auto dis = [&](const char* k, int bit) {
          if (auto* s = f->attr(k)) {
            if (*s == "disable") t.disableflags |= bit;
            else if (*s != "enable") pfail(*f, std::string("flag ") + k + " must be enable/disable");
          }
        };
        dis("constraint", OX_DSBL_CONSTRAINT); dis("limit", OX_DSBL_LIMIT); dis("contact", OX_DSBL_CONTACT);
        dis("passive", OX_DSBL_PASSIVE); dis("gravity", OX_DSBL_GRAVITY); dis("clampctrl", OX_DSBL_CLAMPCTRL);
        dis("warmstart", OX_DSBL_WARMSTART); dis("filterparent", OX_DSBL_FILTERPARENT); dis("equality", OX_DSBL_EQUALITY); dis("frictionloss", OX_DSBL_FRICTIONLOSS);
        dis("actuation", OX_DSBL_ACTUATION); dis("refsafe", OX_DSBL_REFSAFE); dis("eulerdamp", OX_DSBL_EULERDAMP);
        for (const char* k : {"sensor", "autoreset", "spring", "damper"})   // disable flags the step does not implement: refuse, do not ignore
          if (auto* s = f->attr(k))
            if (*s == "disable") cfail(std::string("flag ") + k + "=disable is outside the supported subset");
        for (const char* k : {"energy", "fwdinv", "island", "multiccd", "override", "invdiscrete"})
          if (auto* s = f->attr(k))
            if (*s == "enable") cfail(std::string("flag ") + k + "=enable is outside the supported subset");
      }
    } else if (n == "worldbody") {
      parse_body(B, *ch, 0, "", true);
    } else if (n == "compiler" || n == "default" || n == "size" || n == "visual" || n == "statistic" || n == "custom" ||
               n == "keyframe" || n == "actuator" || n == "sensor" || n == "contact") {
      // handled elsewhere / no effect on the step
      if (n == "statistic" && ch->attr("meaninertia")) cfail("statistic meaninertia (an override of the solver's scale) is outside the supported subset");
    } else if (n == "asset") {
      for (auto& as : ch->children)
        if (as->name == "mesh" || as->name == "hfield") cfail("asset <" + as->name + "> is outside the supported subset");
    } else if (n == "equality") {
      // compiled after the joints and bodies are known (below)
    } else if (n == "tendon") {
      // compiled after the joints are known (below)
    } else if (n == "deformable" || n == "extension") {
      if (!ch->children.empty()) cfail("<" + n + "> is outside the supported subset");
    } else {
      pfail(*ch, "unrecognized top-level element");
    }
  }

  const int nbody = (int)B.bodies.size();
  // ---- body mass properties ----
  for (int i = 1; i < nbody; i++) {
    BodyDef& b = B.bodies[i];
    bool use_geoms = B.c.inertiafromgeom == 1 || (B.c.inertiafromgeom == 2 && !b.has_inertial);
    if (use_geoms) body_inertia_from_geoms(B, b);
    else if (!b.has_inertial) { b.mass = 0; }
    b.mass = std::max(b.mass, B.c.boundmass);   // compiler boundmass / boundinertia: lower bounds for every body but the world
    for (int k = 0; k < 3; k++) b.inertia[k] = std::max(b.inertia[k], B.c.boundinertia);
  }
  if (B.c.settotalmass > 0) {   // mjCModel::SetTotalMass: one scale for all masses and inertias (densities scale, shapes do not)
    double total = 0;
    for (int i = 1; i < nbody; i++) total += B.bodies[i].mass;
    if (total > OX_MINVAL)
      for (int i = 1; i < nbody; i++) {
        B.bodies[i].mass *= B.c.settotalmass / total;
        for (int k = 0; k < 3; k++) B.bodies[i].inertia[k] *= B.c.settotalmass / total;
      }
  }

  // ---- sizes, joint/dof addressing ----
  const int njnt = (int)B.joints.size();
  int nq = 0, nv = 0;
  std::vector<int> jq(njnt), jd(njnt);
  for (int j = 0; j < njnt; j++) {
    jq[j] = nq; jd[j] = nv;
    switch (B.joints[j].type) {
      case OX_JNT_FREE: nq += 7; nv += 6; break;
      case OX_JNT_BALL: nq += 4; nv += 3; break;
      default: nq += 1; nv += 1;
    }
  }
  t.nq = nq; t.nv = nv; t.nbody = nbody; t.njnt = njnt; t.ngeom = (int)B.geoms.size(); t.nsite = (int)B.sites.size();
  t.na = 0;

  auto& nm = M->names;
  nm[OX_OBJ_BODY].resize(nbody);
  nm[OX_OBJ_JOINT].resize(njnt);
  nm[OX_OBJ_GEOM].resize(t.ngeom);
  nm[OX_OBJ_SITE].resize(t.nsite);
  nm[OX_OBJ_ACTUATOR];
  nm[OX_OBJ_SENSOR];

  M->v_body_parentid.assign(nbody, 0); M->v_body_rootid.assign(nbody, 0); M->v_body_weldid.assign(nbody, 0);
  M->v_body_jntadr.assign(nbody, -1); M->v_body_jntnum.assign(nbody, 0);
  M->v_body_dofadr.assign(nbody, -1); M->v_body_dofnum.assign(nbody, 0);
  M->v_body_pos.assign(3 * nbody, 0); M->v_body_quat.assign(4 * nbody, 0);
  M->v_body_ipos.assign(3 * nbody, 0); M->v_body_iquat.assign(4 * nbody, 0);
  M->v_body_mass.assign(nbody, 0); M->v_body_inertia.assign(3 * nbody, 0);
  M->v_body_subtreemass.assign(nbody, 0); M->v_body_invweight0.assign(2 * nbody, 0);
  M->v_body_quat[0] = 1; M->v_body_iquat[0] = 1;

  // joints are already grouped by body in depth-first body order because parse_body adds a body's
  // joints before descending; verify monotonic ownership
  {
    int last = 0;
    for (int j = 0; j < njnt; j++) {
      if (B.joints[j].body < last) cfail("internal: joints not ordered by body");
      last = B.joints[j].body;
    }
  }
  for (int i = 1; i < nbody; i++) {
    const BodyDef& b = B.bodies[i];
    nm[OX_OBJ_BODY][i] = b.name;
    M->v_body_parentid[i] = b.parent;
    M->v_body_rootid[i] = b.parent == 0 ? i : M->v_body_rootid[b.parent];
    M->v_body_weldid[i] = b.joints.empty() ? M->v_body_weldid[b.parent] : i;
    std::memcpy(&M->v_body_pos[3 * i], b.pos, 3 * sizeof(double));
    std::memcpy(&M->v_body_quat[4 * i], b.quat, 4 * sizeof(double));
    std::memcpy(&M->v_body_ipos[3 * i], b.ipos, 3 * sizeof(double));
    std::memcpy(&M->v_body_iquat[4 * i], b.iquat, 4 * sizeof(double));
    M->v_body_mass[i] = b.mass;
    std::memcpy(&M->v_body_inertia[3 * i], b.inertia, 3 * sizeof(double));
    if (!b.joints.empty()) {
      M->v_body_jntadr[i] = b.joints[0];
      M->v_body_jntnum[i] = (int)b.joints.size();
      M->v_body_dofadr[i] = jd[b.joints[0]];
      int nd = 0;
      for (int j : b.joints) {
        int ty = B.joints[j].type;
        nd += ty == OX_JNT_FREE ? 6 : ty == OX_JNT_BALL ? 3 : 1;
        if (ty == OX_JNT_FREE && (b.joints.size() != 1 || b.parent != 0))
          cfail("body '" + b.name + "': a free joint must be the only joint of a child of the world");
      }
      M->v_body_dofnum[i] = nd;
    }
  }
  // inertia-box fluid model (mj_passive -> mj_inertiaBoxFluidModel): every body is replaced by the box with the same mass and
  // principal inertia; the coefficients of its viscous (Stokes, equivalent sphere) and quadratic (face-by-face) drag are constants
  t.nfluid = (t.density > 0 || t.viscosity > 0) ? nbody : 0;
  M->v_body_fluid.assign((size_t)11 * t.nfluid, 0);
  for (int i = 1; i < t.nfluid; i++) {
    const double mass = M->v_body_mass[i], *I = &M->v_body_inertia[3 * i];
    if (mass < OX_MINVAL) continue;   // massless bodies feel no medium
    double bx[3];
    for (int k = 0; k < 3; k++) bx[k] = std::sqrt(std::max(OX_MINVAL, I[(k + 1) % 3] + I[(k + 2) % 3] - I[k]) / mass * 6.0);
    double* f = &M->v_body_fluid[11 * i];
    const double diam = (bx[0] + bx[1] + bx[2]) / 3, pi = 3.14159265358979323846;
    f[0] = pi * diam * diam * diam * t.viscosity;   // torque = -f0 * angular velocity
    f[1] = 3 * pi * diam * t.viscosity;              // force  = -f1 * linear velocity
    for (int k = 0; k < 3; k++) {
      const double a = bx[k], b = bx[(k + 1) % 3], c = bx[(k + 2) % 3];
      f[2 + k] = 0.5 * t.density * b * c;                                   // force_k  -= f * |v_k| v_k
      f[5 + k] = t.density * a * (b * b * b * b + c * c * c * c) / 64;      // torque_k -= f * |w_k| w_k
    }
    for (int k = 0; k < 3; k++) f[8 + k] = t.wind[k];
  }
  t.ngravcomp = 0;
  for (int i = 1; i < nbody; i++) if (B.bodies[i].gravcomp != 0) t.ngravcomp = nbody;
  M->v_body_gravcomp.assign(t.ngravcomp, 0);
  for (int i = 1; i < t.ngravcomp; i++) M->v_body_gravcomp[i] = B.bodies[i].gravcomp;
  nm[OX_OBJ_BODY][0] = "world";
  for (int i = nbody - 1; i >= 0; i--) {
    M->v_body_subtreemass[i] += M->v_body_mass[i];
    if (i > 0) M->v_body_subtreemass[M->v_body_parentid[i]] += M->v_body_subtreemass[i];
  }
  // unique names
  auto check_unique = [&](int ty, const char* what) {
    std::set<std::string> seen;
    for (auto& s : nm[ty])
      if (!s.empty() && !seen.insert(s).second) cfail(std::string("repeated ") + what + " name '" + s + "'");
  };

  M->v_jnt_type.resize(njnt); M->v_jnt_qposadr.resize(njnt); M->v_jnt_dofadr.resize(njnt); M->v_jnt_bodyid.resize(njnt);
  M->v_jnt_limited.resize(njnt);
  M->v_jnt_pos.resize(3 * njnt); M->v_jnt_axis.resize(3 * njnt); M->v_jnt_stiffness.resize(njnt);
  M->v_jnt_range.resize(2 * njnt); M->v_jnt_margin.resize(njnt); M->v_jnt_solref.resize(2 * njnt); M->v_jnt_solimp.resize(5 * njnt);
  M->v_qpos0.assign(nq, 0); M->v_qpos_spring.assign(nq, 0);
  M->v_dof_bodyid.resize(nv); M->v_dof_jntid.resize(nv); M->v_dof_parentid.resize(nv); M->v_dof_Madr.resize(nv); M->v_dof_depth.resize(nv);
  M->v_dof_armature.resize(nv); M->v_dof_damping.resize(nv); M->v_dof_invweight0.assign(nv, 0);
  M->v_dof_frictionloss.assign(nv, 0); M->v_dof_solref_fri.assign(2 * nv, 0); M->v_dof_solimp_fri.assign(5 * nv, 0);
  t.nfloss = 0;
  int nlimited = 0;
  for (int j = 0; j < njnt; j++) {
    const JointDef& jn = B.joints[j];
    nm[OX_OBJ_JOINT][j] = jn.name;
    M->v_jnt_type[j] = jn.type; M->v_jnt_qposadr[j] = jq[j]; M->v_jnt_dofadr[j] = jd[j]; M->v_jnt_bodyid[j] = jn.body;
    M->v_jnt_limited[j] = jn.limited;
    nlimited += jn.limited;
    std::memcpy(&M->v_jnt_pos[3 * j], jn.pos, 3 * sizeof(double));
    std::memcpy(&M->v_jnt_axis[3 * j], jn.axis, 3 * sizeof(double));
    M->v_jnt_stiffness[j] = jn.stiffness;
    M->v_jnt_range[2 * j] = jn.range[0]; M->v_jnt_range[2 * j + 1] = jn.range[1];
    M->v_jnt_margin[j] = jn.margin;
    std::memcpy(&M->v_jnt_solref[2 * j], jn.solref, 2 * sizeof(double));
    std::memcpy(&M->v_jnt_solimp[5 * j], jn.solimp, 5 * sizeof(double));
    int ndof = jn.type == OX_JNT_FREE ? 6 : jn.type == OX_JNT_BALL ? 3 : 1;
    for (int k = 0; k < ndof; k++) {
      int d = jd[j] + k;
      M->v_dof_bodyid[d] = jn.body; M->v_dof_jntid[d] = j;
      M->v_dof_armature[d] = jn.armature; M->v_dof_damping[d] = jn.damping;
      M->v_dof_frictionloss[d] = jn.frictionloss;   // dry friction acts on every dof of the joint
      std::memcpy(&M->v_dof_solref_fri[2 * d], jn.solref_fri, 2 * sizeof(double));
      std::memcpy(&M->v_dof_solimp_fri[5 * d], jn.solimp_fri, 5 * sizeof(double));
      t.nfloss += jn.frictionloss > 0;
    }
    const BodyDef& b = B.bodies[jn.body];
    if (jn.type == OX_JNT_FREE) {
      for (int k = 0; k < 3; k++) M->v_qpos0[jq[j] + k] = b.pos[k];
      for (int k = 0; k < 4; k++) M->v_qpos0[jq[j] + 3 + k] = b.quat[k];
      for (int k = 0; k < 7; k++) M->v_qpos_spring[jq[j] + k] = M->v_qpos0[jq[j] + k];
    } else if (jn.type == OX_JNT_BALL) {
      M->v_qpos0[jq[j]] = 1; M->v_qpos_spring[jq[j]] = 1;
    } else {
      M->v_qpos0[jq[j]] = jn.ref; M->v_qpos_spring[jq[j]] = jn.springref;
    }
  }
  // dof tree
  {
    int nM = 0;
    for (int d = 0; d < nv; d++) {
      int b = M->v_dof_bodyid[d];
      int par;
      if (d > M->v_body_dofadr[b]) par = d - 1;
      else {
        int p = M->v_body_parentid[b];
        while (p > 0 && M->v_body_dofnum[p] == 0) p = M->v_body_parentid[p];
        par = p > 0 ? M->v_body_dofadr[p] + M->v_body_dofnum[p] - 1 : -1;
      }
      M->v_dof_parentid[d] = par;
      M->v_dof_Madr[d] = nM;
      int depth = 1;
      for (int a = par; a >= 0; a = M->v_dof_parentid[a]) depth++;
      M->v_dof_depth[d] = depth;  // number of dofs on the chain d -> root = length of row d of qM
      nM += depth;
    }
    t.nM = nM;
    t.nvv = nv * nv;
    M->v_dof_Mdense.assign((size_t)nv * nv, -1);
    for (int i = 0; i < nv; i++) {
      int adr = M->v_dof_Madr[i];
      for (int j = i; j >= 0; j = M->v_dof_parentid[j], adr++) {
        M->v_dof_Mdense[(size_t)i * nv + j] = adr;
        M->v_dof_Mdense[(size_t)j * nv + i] = adr;
      }
    }
  }

  // ---- geoms, sites ----
  const int ngeom = t.ngeom, nsite = t.nsite;
  // MuJoCo groups geoms by body id; parse order is already body-depth-first with world first, but world
  // geoms declared after child bodies would break it, so sort stably by body.
  std::vector<int> gorder(ngeom);
  for (int i = 0; i < ngeom; i++) gorder[i] = i;
  std::stable_sort(gorder.begin(), gorder.end(), [&](int a, int b) { return B.geoms[a].body < B.geoms[b].body; });
  std::vector<GeomDef> G(ngeom);
  for (int i = 0; i < ngeom; i++) G[i] = B.geoms[gorder[i]];
  std::vector<int> sorder(nsite);
  for (int i = 0; i < nsite; i++) sorder[i] = i;
  std::stable_sort(sorder.begin(), sorder.end(), [&](int a, int b) { return B.sites[a].body < B.sites[b].body; });
  M->v_geom_type.resize(ngeom); M->v_geom_bodyid.resize(ngeom); M->v_geom_contype.resize(ngeom);
  M->v_geom_conaffinity.resize(ngeom); M->v_geom_condim.resize(ngeom); M->v_geom_priority.resize(ngeom);
  M->v_geom_size.resize(3 * ngeom); M->v_geom_pos.resize(3 * ngeom); M->v_geom_quat.resize(4 * ngeom);
  M->v_geom_friction.resize(3 * ngeom); M->v_geom_solmix.resize(ngeom); M->v_geom_solref.resize(2 * ngeom);
  M->v_geom_solimp.resize(5 * ngeom); M->v_geom_margin.resize(ngeom); M->v_geom_gap.resize(ngeom);
  for (int i = 0; i < ngeom; i++) {
    const GeomDef& g = G[i];
    nm[OX_OBJ_GEOM][i] = g.name;
    M->v_geom_type[i] = g.type; M->v_geom_bodyid[i] = g.body; M->v_geom_contype[i] = g.contype;
    M->v_geom_conaffinity[i] = g.conaffinity; M->v_geom_condim[i] = g.condim; M->v_geom_priority[i] = g.priority;
    std::memcpy(&M->v_geom_size[3 * i], g.size, 3 * sizeof(double));
    std::memcpy(&M->v_geom_pos[3 * i], g.pos, 3 * sizeof(double));
    std::memcpy(&M->v_geom_quat[4 * i], g.quat, 4 * sizeof(double));
    std::memcpy(&M->v_geom_friction[3 * i], g.friction, 3 * sizeof(double));
    M->v_geom_solmix[i] = g.solmix;
    std::memcpy(&M->v_geom_solref[2 * i], g.solref, 2 * sizeof(double));
    std::memcpy(&M->v_geom_solimp[5 * i], g.solimp, 5 * sizeof(double));
    M->v_geom_margin[i] = g.margin; M->v_geom_gap[i] = g.gap;
  }
  M->v_site_bodyid.resize(nsite); M->v_site_pos.resize(3 * nsite); M->v_site_quat.resize(4 * nsite);
  M->v_site_type.resize(nsite); M->v_site_size.resize(3 * nsite);
  for (int i = 0; i < nsite; i++) {
    const SiteDef& s = B.sites[sorder[i]];
    nm[OX_OBJ_SITE][i] = s.name;
    M->v_site_bodyid[i] = s.body;
    std::memcpy(&M->v_site_pos[3 * i], s.pos, 3 * sizeof(double));
    std::memcpy(&M->v_site_quat[4 * i], s.quat, 4 * sizeof(double));
    M->v_site_type[i] = s.type;
    std::memcpy(&M->v_site_size[3 * i], s.size, 3 * sizeof(double));
  }
  check_unique(OX_OBJ_BODY, "body"); check_unique(OX_OBJ_JOINT, "joint");
  check_unique(OX_OBJ_GEOM, "geom"); check_unique(OX_OBJ_SITE, "site");

  // ---- contact excludes ----
  for (auto& ch : root->children)
    if (ch->name == "contact")
      for (auto& e : ch->children) {
        if (e->name == "exclude") {
          check_attrs(*e, "exclude");
          auto* b1 = e->attr("body1");
          auto* b2 = e->attr("body2");
          if (!b1 || !b2) pfail(*e, "exclude requires body1 and body2");
          int i1 = find_name(nm[OX_OBJ_BODY], *b1), i2 = find_name(nm[OX_OBJ_BODY], *b2);
          if (i1 < 0 || i2 < 0) cfail("exclude: unknown body '" + (i1 < 0 ? *b1 : *b2) + "'");
          B.excludes.insert({std::min(i1, i2), std::max(i1, i2)});
        } else cfail("<contact><" + e->name + "> is outside the supported subset");
      }

  // ---- candidate collision pairs (broadphase filters applied once) ----
  {
    const bool contact_on = !(t.disableflags & (OX_DSBL_CONTACT | OX_DSBL_CONSTRAINT));
    const bool filterparent = !(t.disableflags & OX_DSBL_FILTERPARENT);
    int nconmax = 0, ncontact_rows = 0;
    for (int b1 = 0; b1 < nbody && contact_on; b1++)
      for (int b2 = b1 + 1; b2 < nbody; b2++) {
        int w1 = M->v_body_weldid[b1], w2 = M->v_body_weldid[b2];
        if (w1 == w2) continue;  // same weld body (covers both static)
        int wp1 = M->v_body_weldid[M->v_body_parentid[w1]], wp2 = M->v_body_weldid[M->v_body_parentid[w2]];
        if (filterparent && w1 != 0 && w2 != 0 && (w1 == wp2 || w2 == wp1)) continue;
        if (B.excludes.count({b1, b2})) continue;
        for (int g1 = 0; g1 < ngeom; g1++) {
          if (G[g1].body != b1) continue;
          for (int g2 = 0; g2 < ngeom; g2++) {
            if (G[g2].body != b2) continue;
            const GeomDef &A = G[g1], &Bg = G[g2];
            if (!((A.contype & Bg.conaffinity) || (Bg.contype & A.conaffinity))) continue;
            int ga = g1, gb = g2;
            if (G[ga].type > G[gb].type) std::swap(ga, gb);
            int ta = G[ga].type, tb = G[gb].type;
            int maxcon = 0;
            if (ta == OX_GEOM_PLANE && tb == OX_GEOM_PLANE) continue;
            if (ta == OX_GEOM_PLANE && tb == OX_GEOM_SPHERE) maxcon = 1;
            else if (ta == OX_GEOM_PLANE && tb == OX_GEOM_CAPSULE) maxcon = 2;
            else if (ta == OX_GEOM_PLANE && tb == OX_GEOM_BOX) maxcon = 4;
            else if (ta == OX_GEOM_SPHERE && tb == OX_GEOM_SPHERE) maxcon = 1;
            else if (ta == OX_GEOM_SPHERE && tb == OX_GEOM_CAPSULE) maxcon = 1;
            else if (ta == OX_GEOM_CAPSULE && tb == OX_GEOM_CAPSULE) maxcon = 2;
            else if (ta == OX_GEOM_SPHERE && tb == OX_GEOM_BOX) maxcon = 1;
            else if (ta == OX_GEOM_CAPSULE && tb == OX_GEOM_BOX) maxcon = 2;
            else if (ta == OX_GEOM_BOX && tb == OX_GEOM_BOX) maxcon = 8;
            else {
              static const char* tn[] = {"plane", "hfield", "sphere", "capsule", "ellipsoid", "cylinder", "box", "mesh"};
              cfail(std::string("collision pair ") + tn[ta] + "-" + tn[tb] + " (geoms '" + G[ga].name + "', '" + G[gb].name +
                    "') is outside the supported narrowphase set; mask it with contype/conaffinity or <exclude>");
            }
            const GeomDef &P = G[ga], &Q = G[gb];
            int dim;
            double fri[3], solref[2], solimp[5];
            if (P.priority != Q.priority) {
              const GeomDef& H = P.priority > Q.priority ? P : Q;
              dim = H.condim;
              std::memcpy(fri, H.friction, sizeof fri);
              std::memcpy(solref, H.solref, sizeof solref);
              std::memcpy(solimp, H.solimp, sizeof solimp);
            } else {
              dim = std::max(P.condim, Q.condim);
              double mix;
              if (P.solmix >= OX_MINVAL && Q.solmix >= OX_MINVAL) mix = P.solmix / (P.solmix + Q.solmix);
              else if (P.solmix < OX_MINVAL && Q.solmix < OX_MINVAL) mix = 0.5;
              else if (P.solmix < OX_MINVAL) mix = 0;
              else mix = 1;
              if (P.solref[0] > 0 && Q.solref[0] > 0)
                for (int k = 0; k < 2; k++) solref[k] = mix * P.solref[k] + (1 - mix) * Q.solref[k];
              else
                for (int k = 0; k < 2; k++) solref[k] = std::min(P.solref[k], Q.solref[k]);
              for (int k = 0; k < 5; k++) solimp[k] = mix * P.solimp[k] + (1 - mix) * Q.solimp[k];
              for (int k = 0; k < 3; k++) fri[k] = std::max(P.friction[k], Q.friction[k]);
            }
            if (dim != 1 && dim != 3 && dim != 4 && dim != 6) cfail("contact dimension " + std::to_string(dim) + " is not a valid condim (1, 3, 4 or 6)");
            M->v_pair_geom1.push_back(ga); M->v_pair_geom2.push_back(gb);
            M->v_pair_dim.push_back(dim); M->v_pair_maxcon.push_back(maxcon); M->v_pair_conadr.push_back(nconmax);
            double f5[5] = {fri[0], fri[0], fri[1], fri[2], fri[2]};
            for (double v : f5) M->v_pair_friction.push_back(v);
            for (double v : solref) M->v_pair_solref.push_back(v);
            for (double v : solimp) M->v_pair_solimp.push_back(v);
            M->v_pair_margin.push_back(std::max(P.margin, Q.margin));
            M->v_pair_gap.push_back(std::max(P.gap, Q.gap));
            nconmax += maxcon;
            ncontact_rows += maxcon * (dim == 1 ? 1 : 2 * (dim - 1));
          }
        }
      }
    t.npair = (int)M->v_pair_geom1.size();
    t.nconmax = nconmax;
    bool limits_on = !(t.disableflags & (OX_DSBL_LIMIT | OX_DSBL_CONSTRAINT));
    t.nefcmax = (limits_on ? 2 * nlimited : 0) + ncontact_rows + ((t.disableflags & OX_DSBL_CONSTRAINT) ? 0 : t.nfloss);   // equality rows are added once they are compiled (below)
  }

  // ---- fixed tendons: length = sum_i coef_i * q_i over hinge / slide joints; limits, spring (with dead band) and damper.
  // Spatial tendons (site / geom wrapping), tendon friction loss and tendon transmissions are outside the supported subset.
  t.ntendon = 0; t.nwrap = 0;
  std::vector<char> spring_default;   // springlength unspecified: the length at qpos0 (known for spatial tendons only after the qpos0 pass)
  for (auto& ch : root->children)
    if (ch->name == "tendon")
      for (auto& e : ch->children) {
        if (e->name != "fixed" && e->name != "spatial") cfail("tendon <" + e->name + "> is outside the supported subset (fixed, spatial)");
        const bool spatial = e->name == "spatial";
        check_attrs(*e, "tendon_fixed");
        Attrs a = merged(B.c, *e, "tendon", "");
        const std::string tname = a.str_or("name", "");
        if (a.num("frictionloss", 0) != 0) cfail("tendon '" + tname + "': frictionloss is outside the supported subset");
        double range[2] = {0, 0}, solref[2] = {0.02, 1}, solimp[5] = {0.9, 0.95, 0.001, 0.5, 2}, spring[2] = {-1, -1};
        const bool has_range = a.has("range");
        a.vec("range", range, 2);
        a.vec("solreflimit", solref, 2);
        a.vec("solimplimit", solimp, 5, true);
        if (a.has("springlength")) { const auto sv = a.nums("springlength"); a.vec("springlength", spring, 2, true); if (sv.size() == 1) spring[1] = spring[0]; }
        const int lim = a.boolean("limited");
        if (lim < 0 && has_range && !B.c.autolimits) cfail("tendon '" + tname + "': range given but limited unspecified and autolimits=false");
        const bool limited = lim >= 0 ? lim == 1 : has_range;
        if (limited && !(range[0] < range[1])) cfail("tendon '" + tname + "': limited tendon needs range[0] < range[1]");
        M->v_tendon_adr.push_back(t.nwrap);
        int num = 0;
        double len0 = 0;
        for (auto& w : e->children) {
          if (spatial) {   // straight segments through sites; wrapping geoms and pulleys are outside the supported subset
            if (w->name != "site") cfail("tendon '" + tname + "': <" + w->name + "> in a spatial tendon is outside the supported subset (site)");
            check_attrs(*w, "tendon_site");
            const std::string* sn = w->attr("site");
            if (!sn) pfail(*w, "tendon <site> requires site");
            const int sid = find_name(nm[OX_OBJ_SITE], *sn);
            if (sid < 0) cfail("tendon '" + tname + "': unknown site '" + *sn + "'");
            M->v_wrap_objid.push_back(sid); M->v_wrap_prm.push_back(0);
            num++; t.nwrap++;
            continue;
          }
          if (w->name != "joint") pfail(*w, "a fixed tendon holds <joint> elements only");
          check_attrs(*w, "tendon_joint");
          const std::string* jn = w->attr("joint");
          const std::string* cf = w->attr("coef");
          if (!jn || !cf) pfail(*w, "tendon <joint> requires joint and coef");
          const int j = find_name(nm[OX_OBJ_JOINT], *jn);
          if (j < 0) cfail("tendon '" + tname + "': unknown joint '" + *jn + "'");
          if (B.joints[j].type != OX_JNT_HINGE && B.joints[j].type != OX_JNT_SLIDE) cfail("tendon '" + tname + "': fixed tendons combine hinge / slide joints");
          const double coef = std::stod(*cf);
          M->v_wrap_objid.push_back(j); M->v_wrap_prm.push_back(coef);
          len0 += coef * M->v_qpos0[M->v_jnt_qposadr[j]];
          num++; t.nwrap++;
        }
        if (!num) cfail("tendon '" + tname + "': no joints");
        if (spatial && num < 2) cfail("tendon '" + tname + "': a spatial tendon needs at least two sites");
        M->v_tendon_type.push_back(spatial ? OX_TEN_SPATIAL : OX_TEN_FIXED);
        spring_default.push_back(spring[0] < 0 && spring[1] < 0);
        M->v_tendon_num.push_back(num); M->v_tendon_limited.push_back(limited ? 1 : 0);
        M->v_tendon_range.push_back(range[0]); M->v_tendon_range.push_back(range[1]);
        M->v_tendon_margin.push_back(a.num("margin", 0));
        for (double v : solref) M->v_tendon_solref_lim.push_back(v);
        for (double v : solimp) M->v_tendon_solimp_lim.push_back(v);
        M->v_tendon_stiffness.push_back(a.num("stiffness", 0)); M->v_tendon_damping.push_back(a.num("damping", 0));
        if (a.num("damping", 0) != 0 && t.integrator == OX_INT_IMPLICITFAST)
          cfail("tendon '" + tname + "': tendon damping with the implicitfast integrator (off-diagonal velocity derivative) is outside the supported subset");
        // springlength -1 = "use the length at qpos0" (mj_setLengthRange convention of the compiler)
        M->v_tendon_lengthspring.push_back(spring[0] < 0 && spring[1] < 0 ? len0 : spring[0]);
        M->v_tendon_lengthspring.push_back(spring[0] < 0 && spring[1] < 0 ? len0 : spring[1]);
        M->v_tendon_length0.push_back(len0); M->v_tendon_invweight0.push_back(0);
        nm[OX_OBJ_TENDON].push_back(tname);
        if (limited && !(t.disableflags & (OX_DSBL_LIMIT | OX_DSBL_CONSTRAINT))) t.nefcmax += 2;
        t.ntendon++;
      }
  check_unique(OX_OBJ_TENDON, "tendon");

  // ---- actuators ----
  std::vector<ActDef> acts;
  for (auto& ch : root->children)
    if (ch->name == "actuator")
      for (auto& e : ch->children) {
        if (!is_actuator_tag(e->name)) cfail("actuator <" + e->name + "> is outside the supported subset (motor, position, velocity, general)");
        check_attrs(*e, "actuator_common");
        Attrs a = merged(B.c, *e, "actuator", "");
        ActDef ad;
        ad.name = a.str_or("name", "");
        for (const char* k : {"site", "body", "jointinparent", "slidersite", "cranksite"})
          if (a.has(k)) cfail("actuator '" + ad.name + "': transmission '" + k + "' is outside the supported subset (joint, tendon)");
        if (a.has("joint") == a.has("tendon")) pfail(*e, "actuator requires exactly one of joint / tendon");
        if (a.has("tendon")) {   // fixed-tendon transmission: length = gear * tendon length, moment = gear * tendon Jacobian
          ad.tendon = find_name(nm[OX_OBJ_TENDON], a.str("tendon"));
          if (ad.tendon < 0) cfail("actuator '" + ad.name + "': unknown tendon '" + a.str("tendon") + "'");
          if (t.integrator == OX_INT_IMPLICITFAST) cfail("actuator '" + ad.name + "': tendon transmissions with the implicitfast integrator (off-diagonal velocity derivative) are outside the supported subset");
        } else {
          ad.joint = find_name(nm[OX_OBJ_JOINT], a.str("joint"));
          if (ad.joint < 0) cfail("actuator '" + ad.name + "': unknown joint '" + a.str("joint") + "'");
          int jt = B.joints[ad.joint].type;
          if (jt != OX_JNT_HINGE && jt != OX_JNT_SLIDE) cfail("actuator '" + ad.name + "': only hinge/slide joint transmissions are supported");
        }
        if (a.has("gear")) ad.gear = a.nums("gear").at(0);
        if (a.has("dyntype")) {
          const std::string& s = a.str("dyntype");
          if (e->name != "general" && s != "none") cfail("actuator '" + ad.name + "': dyntype is an attribute of <general>");
          if (s == "none") ad.dyntype = OX_DYN_NONE;
          else if (s == "integrator") ad.dyntype = OX_DYN_INTEGRATOR;
          else if (s == "filter") ad.dyntype = OX_DYN_FILTER;
          else if (s == "filterexact") ad.dyntype = OX_DYN_FILTEREXACT;
          else cfail("actuator '" + ad.name + "': dyntype '" + s + "' is outside the supported subset (none, integrator, filter, filterexact)");
        }
        if (a.has("inheritrange") && a.num("inheritrange", 0) != 0) cfail("actuator '" + ad.name + "': inheritrange is outside the supported subset (give ctrlrange)");
        if (a.has("actdim") || a.has("actearly")) cfail("actuator '" + ad.name + "': actdim / actearly are outside the supported subset");
        a.vec("dynprm", ad.dynprm, 3, true);
        if (e->name == "position" && a.num("timeconst", 0) > 0) {   // <position timeconst>: first-order filter on the target, exact integration
          ad.dyntype = OX_DYN_FILTEREXACT;
          ad.dynprm[0] = a.num("timeconst", 0);
        }
        {
          const bool has_ar = a.has("actrange");
          if (has_ar) a.vec("actrange", ad.actrange, 2);
          const int al = a.boolean("actlimited");
          ad.actlimited = al >= 0 ? al : (has_ar && B.c.autolimits);
          if (ad.actlimited && ad.dyntype == OX_DYN_NONE) cfail("actuator '" + ad.name + "': actlimited needs a stateful actuator (dyntype)");
          if (ad.actlimited && !(ad.actrange[0] < ad.actrange[1])) cfail("actuator '" + ad.name + "': invalid actrange");
        }
        if (e->name == "position") {
          double kp = a.num("kp", 1), kv = a.num("kv", 0);
          if (a.num("timeconst", 0) < 0) cfail("actuator '" + ad.name + "': timeconst must be >= 0");
          if (a.has("dampratio")) cfail("actuator '" + ad.name + "': dampratio is outside the supported subset (give kv)");
          ad.gainprm[0] = kp; ad.biastype = OX_BIAS_AFFINE; ad.biasprm[1] = -kp; ad.biasprm[2] = -kv;
        } else if (e->name == "velocity") {
          double kv = a.num("kv", 1);
          ad.gainprm[0] = kv; ad.biastype = OX_BIAS_AFFINE; ad.biasprm[2] = -kv;
        } else if (e->name == "general") {
          if (a.has("gaintype")) {
            const std::string& s = a.str("gaintype");
            if (s == "fixed") ad.gaintype = OX_GAIN_FIXED;
            else if (s == "affine") ad.gaintype = OX_GAIN_AFFINE;
            else cfail("actuator '" + ad.name + "': gaintype '" + s + "' unsupported");
          }
          if (a.has("biastype")) {
            const std::string& s = a.str("biastype");
            if (s == "none") ad.biastype = OX_BIAS_NONE;
            else if (s == "affine") ad.biastype = OX_BIAS_AFFINE;
            else cfail("actuator '" + ad.name + "': biastype '" + s + "' unsupported");
          }
          a.vec("gainprm", ad.gainprm, 3, true);
          a.vec("biasprm", ad.biasprm, 3, true);
        }
        bool has_cr = a.has("ctrlrange"), has_fr = a.has("forcerange");
        if (has_cr) a.vec("ctrlrange", ad.ctrlrange, 2);
        if (has_fr) a.vec("forcerange", ad.forcerange, 2);
        int cl = a.boolean("ctrllimited"), fl = a.boolean("forcelimited");
        ad.ctrllimited = cl >= 0 ? cl : (has_cr && B.c.autolimits);
        ad.forcelimited = fl >= 0 ? fl : (has_fr && B.c.autolimits);
        if (ad.ctrllimited && !(ad.ctrlrange[0] < ad.ctrlrange[1])) cfail("actuator '" + ad.name + "': invalid ctrlrange");
        if (ad.forcelimited && !(ad.forcerange[0] < ad.forcerange[1])) cfail("actuator '" + ad.name + "': invalid forcerange");
        acts.push_back(ad);
      }
  const int nu = (int)acts.size();
  t.nu = nu;
  M->v_actuator_trnid.resize(nu); M->v_actuator_trntype.resize(nu); M->v_actuator_gaintype.resize(nu); M->v_actuator_biastype.resize(nu);
  M->v_actuator_ctrllimited.resize(nu); M->v_actuator_forcelimited.resize(nu);
  M->v_actuator_gear.resize(nu); M->v_actuator_gainprm.resize(3 * nu); M->v_actuator_biasprm.resize(3 * nu);
  M->v_actuator_ctrlrange.resize(2 * nu); M->v_actuator_forcerange.resize(2 * nu);
  M->v_actuator_dyntype.resize(nu); M->v_actuator_actadr.resize(nu); M->v_actuator_actlimited.resize(nu);
  M->v_actuator_dynprm.resize(3 * nu); M->v_actuator_actrange.resize(2 * nu);
  nm[OX_OBJ_ACTUATOR].resize(nu);
  int na = 0;
  for (int i = 0; i < nu; i++) {
    const ActDef& a = acts[i];
    nm[OX_OBJ_ACTUATOR][i] = a.name;
    M->v_actuator_trnid[i] = a.tendon >= 0 ? a.tendon : a.joint; M->v_actuator_trntype[i] = a.tendon >= 0 ? OX_TRN_TENDON : OX_TRN_JOINT;
    M->v_actuator_gaintype[i] = a.gaintype; M->v_actuator_biastype[i] = a.biastype;
    M->v_actuator_ctrllimited[i] = a.ctrllimited; M->v_actuator_forcelimited[i] = a.forcelimited;
    M->v_actuator_gear[i] = a.gear;
    for (int k = 0; k < 3; k++) { M->v_actuator_gainprm[3 * i + k] = a.gainprm[k]; M->v_actuator_biasprm[3 * i + k] = a.biasprm[k]; }
    for (int k = 0; k < 2; k++) { M->v_actuator_ctrlrange[2 * i + k] = a.ctrlrange[k]; M->v_actuator_forcerange[2 * i + k] = a.forcerange[k]; }
    M->v_actuator_dyntype[i] = a.dyntype; M->v_actuator_actlimited[i] = a.actlimited;
    M->v_actuator_actadr[i] = a.dyntype != OX_DYN_NONE ? na++ : -1;   // one activation variable per stateful actuator (actdim 1)
    for (int k = 0; k < 3; k++) M->v_actuator_dynprm[3 * i + k] = a.dynprm[k];
    for (int k = 0; k < 2; k++) M->v_actuator_actrange[2 * i + k] = a.actrange[k];
  }
  t.na = na;
  check_unique(OX_OBJ_ACTUATOR, "actuator");

  // ---- sensors (N2 subset) ----
  {
    int adr = 0;
    for (auto& ch : root->children)
      if (ch->name == "sensor")
        for (auto& e : ch->children) {
          check_attrs(*e, "sensor_common");
          struct Spec { const char* tag; int type; int objtype; const char* attr; int dim; };
          static const Spec specs[] = {
              {"jointpos", OX_SENS_JOINTPOS, OX_OBJ_JOINT, "joint", 1}, {"jointvel", OX_SENS_JOINTVEL, OX_OBJ_JOINT, "joint", 1},
              {"actuatorpos", OX_SENS_ACTUATORPOS, OX_OBJ_ACTUATOR, "actuator", 1},
              {"actuatorvel", OX_SENS_ACTUATORVEL, OX_OBJ_ACTUATOR, "actuator", 1},
              {"actuatorfrc", OX_SENS_ACTUATORFRC, OX_OBJ_ACTUATOR, "actuator", 1},
              {"subtreecom", OX_SENS_SUBTREECOM, OX_OBJ_BODY, "body", 3},
              {"subtreelinvel", OX_SENS_SUBTREELINVEL, OX_OBJ_BODY, "body", 3},
              {"velocimeter", OX_SENS_VELOCIMETER, OX_OBJ_SITE, "site", 3}, {"gyro", OX_SENS_GYRO, OX_OBJ_SITE, "site", 3},
              {"accelerometer", OX_SENS_ACCELEROMETER, OX_OBJ_SITE, "site", 3}, {"touch", OX_SENS_TOUCH, OX_OBJ_SITE, "site", 1},
              {"force", OX_SENS_FORCE, OX_OBJ_SITE, "site", 3}, {"torque", OX_SENS_TORQUE, OX_OBJ_SITE, "site", 3},
              {"framepos", OX_SENS_FRAMEPOS, -1, "objname", 3}, {"framequat", OX_SENS_FRAMEQUAT, -1, "objname", 4},
              {"framelinvel", OX_SENS_FRAMELINVEL, -1, "objname", 3}, {"frameangvel", OX_SENS_FRAMEANGVEL, -1, "objname", 3},
              {"tendonpos", OX_SENS_TENDONPOS, OX_OBJ_TENDON, "tendon", 1}, {"tendonvel", OX_SENS_TENDONVEL, OX_OBJ_TENDON, "tendon", 1},
              {"framexaxis", OX_SENS_FRAMEXAXIS, -1, "objname", 3}, {"frameyaxis", OX_SENS_FRAMEYAXIS, -1, "objname", 3},
              {"framezaxis", OX_SENS_FRAMEZAXIS, -1, "objname", 3},
              {"framelinacc", OX_SENS_FRAMELINACC, -1, "objname", 3}, {"frameangacc", OX_SENS_FRAMEANGACC, -1, "objname", 3},
              {"ballquat", OX_SENS_BALLQUAT, OX_OBJ_JOINT, "joint", 4}, {"ballangvel", OX_SENS_BALLANGVEL, OX_OBJ_JOINT, "joint", 3},
              {"jointactuatorfrc", OX_SENS_JOINTACTFRC, OX_OBJ_JOINT, "joint", 1},
              {"clock", OX_SENS_CLOCK, OX_OBJ_UNKNOWN, nullptr, 1},
          };
          const Spec* sp = nullptr;
          for (auto& s : specs)
            if (e->name == s.tag) sp = &s;
          if (!sp) cfail("sensor <" + e->name + "> is outside the supported subset");
          if (e->attr("reftype") || e->attr("refname")) cfail("sensor reference frames (reftype/refname) are outside the supported subset");
          if (auto* c = e->attr("cutoff")) if (std::stod(*c) > 0) cfail("sensor cutoff is outside the supported subset");
          int objtype = sp->objtype, objid = -1;
          if (sp->attr) {
            const std::string* on = e->attr(sp->attr);
            if (!on) pfail(*e, std::string("sensor requires attribute '") + sp->attr + "'");
            if (objtype < 0) {
              const std::string* ot = e->attr("objtype");
              if (!ot) pfail(*e, "frame sensor requires objtype");
              if (*ot == "body") objtype = OX_OBJ_BODY;
              else if (*ot == "xbody") objtype = OX_OBJ_XBODY;
              else if (*ot == "geom") objtype = OX_OBJ_GEOM;
              else if (*ot == "site") objtype = OX_OBJ_SITE;
              else cfail("frame sensor objtype '" + *ot + "' is outside the supported subset");
            }
            int ntype = objtype == OX_OBJ_XBODY ? OX_OBJ_BODY : objtype;
            objid = find_name(nm[ntype], *on);
            if (objid < 0) cfail("sensor: unknown object '" + *on + "'");
            if ((sp->type == OX_SENS_JOINTPOS || sp->type == OX_SENS_JOINTVEL) &&
                B.joints[objid].type != OX_JNT_HINGE && B.joints[objid].type != OX_JNT_SLIDE)
              cfail("jointpos/jointvel sensors require a hinge or slide joint");
            if (sp->type == OX_SENS_JOINTACTFRC && B.joints[objid].type != OX_JNT_HINGE && B.joints[objid].type != OX_JNT_SLIDE)
              cfail("jointactuatorfrc sensors require a hinge or slide joint");
            if ((sp->type == OX_SENS_BALLQUAT || sp->type == OX_SENS_BALLANGVEL) && B.joints[objid].type != OX_JNT_BALL)
              cfail("ballquat/ballangvel sensors require a ball joint");
            if (sp->type == OX_SENS_TOUCH) {
              const int st = M->v_site_type[objid];
              if (st != OX_GEOM_SPHERE && st != OX_GEOM_CAPSULE && st != OX_GEOM_BOX)
                cfail("touch sensor '" + *on + "': sensing volumes are sphere, capsule or box sites (ellipsoid / cylinder sites are outside the supported subset)");
            }
          }
          M->v_sensor_type.push_back(sp->type); M->v_sensor_objtype.push_back(objtype); M->v_sensor_objid.push_back(objid);
          M->v_sensor_adr.push_back(adr); M->v_sensor_dim.push_back(sp->dim);
          adr += sp->dim;
          nm[OX_OBJ_SENSOR].push_back(e->attr("name") ? *e->attr("name") : "");
        }
    t.nsensor = (int)M->v_sensor_type.size();
    t.nsensordata = adr;
  }

  // ---- quantities at qpos0: meaninertia, dof_invweight0, body_invweight0 ----
  // Dense, algorithm-independent of the device CRB path: M = sum_b Jb^T diag(m, I_b) Jb + armature.
  if (nv > 0) {
    std::vector<double> xpos(3 * nbody, 0), xquat(4 * nbody, 0), xipos(3 * nbody, 0), ximat(9 * nbody, 0);
    xquat[0] = 1;
    for (int i = 1; i < nbody; i++) {
      int p = M->v_body_parentid[i];
      double r[3];
      hm::rotvec(r, &M->v_body_pos[3 * i], &xquat[4 * p]);
      for (int k = 0; k < 3; k++) xpos[3 * i + k] = xpos[3 * p + k] + r[k];
      hm::mulquat(&xquat[4 * i], &xquat[4 * p], &M->v_body_quat[4 * i]);
      hm::normalize4(&xquat[4 * i]);
      hm::rotvec(r, &M->v_body_ipos[3 * i], &xquat[4 * i]);
      for (int k = 0; k < 3; k++) xipos[3 * i + k] = xpos[3 * i + k] + r[k];
      double qi[4];
      hm::mulquat(qi, &xquat[4 * i], &M->v_body_iquat[4 * i]);
      hm::quat2mat(&ximat[9 * i], qi);
    }
    // Jacobian of (point on body b) : jp[3][nv], jr[3][nv]
    auto jac = [&](int b, const double* pt, std::vector<double>& jp, std::vector<double>& jr) {
      jp.assign(3 * nv, 0); jr.assign(3 * nv, 0);
      while (b > 0 && M->v_body_dofnum[b] == 0) b = M->v_body_parentid[b];
      if (b == 0) return;
      for (int d = M->v_body_dofadr[b] + M->v_body_dofnum[b] - 1; d >= 0; d = M->v_dof_parentid[d]) {
        int j = M->v_dof_jntid[d], jb = M->v_jnt_bodyid[j], ty = M->v_jnt_type[j], k = d - M->v_jnt_dofadr[j];
        double R[9], anchor[3], ax[3], rj[3];
        hm::quat2mat(R, &xquat[4 * jb]);
        hm::rotvec(rj, &M->v_jnt_pos[3 * j], &xquat[4 * jb]);
        for (int c = 0; c < 3; c++) anchor[c] = xpos[3 * jb + c] + rj[c];
        bool rot = true;
        if (ty == OX_JNT_HINGE) hm::rotvec(ax, &M->v_jnt_axis[3 * j], &xquat[4 * jb]);
        else if (ty == OX_JNT_SLIDE) { hm::rotvec(ax, &M->v_jnt_axis[3 * j], &xquat[4 * jb]); rot = false; }
        else if (ty == OX_JNT_BALL) { ax[0] = R[k]; ax[1] = R[3 + k]; ax[2] = R[6 + k]; }
        else {  // free
          if (k < 3) { ax[0] = ax[1] = ax[2] = 0; ax[k] = 1; rot = false; }
          else { ax[0] = R[k - 3]; ax[1] = R[3 + k - 3]; ax[2] = R[6 + k - 3]; }
          for (int c = 0; c < 3; c++) anchor[c] = xpos[3 * jb + c];
        }
        if (rot) {
          double off[3] = {pt[0] - anchor[0], pt[1] - anchor[1], pt[2] - anchor[2]}, cr[3];
          hm::cross(cr, ax, off);
          for (int c = 0; c < 3; c++) { jr[c * nv + d] = ax[c]; jp[c * nv + d] = cr[c]; }
        } else
          for (int c = 0; c < 3; c++) jp[c * nv + d] = ax[c];
      }
    };
    std::vector<double> Md(nv * nv, 0), jp, jr;
    for (int b = 1; b < nbody; b++) {
      double m = M->v_body_mass[b];
      if (m == 0 && M->v_body_inertia[3 * b] == 0) continue;
      jac(b, &xipos[3 * b], jp, jr);
      const double* R = &ximat[9 * b];
      const double* I = &M->v_body_inertia[3 * b];
      // W = R diag(I) R^T
      double W[9];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
          double v = 0;
          for (int k = 0; k < 3; k++) v += R[3 * r + k] * I[k] * R[3 * c + k];
          W[3 * r + c] = v;
        }
      for (int i = 0; i < nv; i++)
        for (int j = 0; j < nv; j++) {
          double v = 0;
          for (int c = 0; c < 3; c++) v += m * jp[c * nv + i] * jp[c * nv + j];
          for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) v += jr[r * nv + i] * W[3 * r + c] * jr[c * nv + j];
          Md[i * nv + j] += v;
        }
    }
    for (int i = 0; i < nv; i++) Md[i * nv + i] += M->v_dof_armature[i];
    double mean = 0;
    for (int i = 0; i < nv; i++) mean += Md[i * nv + i];
    t.meaninertia = mean / nv;
    // Cholesky M = L L^T, then Minv
    std::vector<double> L(Md);
    for (int j = 0; j < nv; j++) {
      double s = L[j * nv + j];
      for (int k = 0; k < j; k++) s -= L[j * nv + k] * L[j * nv + k];
      if (!(s > OX_MINVAL)) cfail("mass matrix at qpos0 is not positive definite (massless moving body?)");
      L[j * nv + j] = std::sqrt(s);
      for (int i = j + 1; i < nv; i++) {
        double v = L[i * nv + j];
        for (int k = 0; k < j; k++) v -= L[i * nv + k] * L[j * nv + k];
        L[i * nv + j] = v / L[j * nv + j];
      }
    }
    std::vector<double> Minv(nv * nv, 0);
    for (int c = 0; c < nv; c++) {
      std::vector<double> x(nv, 0);
      x[c] = 1;
      for (int i = 0; i < nv; i++) {
        double v = x[i];
        for (int k = 0; k < i; k++) v -= L[i * nv + k] * x[k];
        x[i] = v / L[i * nv + i];
      }
      for (int i = nv - 1; i >= 0; i--) {
        double v = x[i];
        for (int k = i + 1; k < nv; k++) v -= L[k * nv + i] * x[k];
        x[i] = v / L[i * nv + i];
      }
      for (int i = 0; i < nv; i++) Minv[i * nv + c] = x[i];
    }
    for (int j = 0; j < njnt; j++) {
      int d = jd[j], ty = M->v_jnt_type[j];
      if (ty == OX_JNT_FREE) {
        double a = (Minv[d * nv + d] + Minv[(d + 1) * nv + d + 1] + Minv[(d + 2) * nv + d + 2]) / 3;
        double r = (Minv[(d + 3) * nv + d + 3] + Minv[(d + 4) * nv + d + 4] + Minv[(d + 5) * nv + d + 5]) / 3;
        for (int k = 0; k < 3; k++) { M->v_dof_invweight0[d + k] = a; M->v_dof_invweight0[d + 3 + k] = r; }
      } else if (ty == OX_JNT_BALL) {
        double a = (Minv[d * nv + d] + Minv[(d + 1) * nv + d + 1] + Minv[(d + 2) * nv + d + 2]) / 3;
        for (int k = 0; k < 3; k++) M->v_dof_invweight0[d + k] = a;
      } else M->v_dof_invweight0[d] = Minv[d * nv + d];
    }
    for (int i = 0; i < t.ntendon; i++) {   // tendon_invweight0 = J M^-1 J' with the tendon Jacobian at qpos0 (constant for fixed tendons)
      std::vector<double> J0(nv, 0.0);
      const int adr = M->v_tendon_adr[i], num = M->v_tendon_num[i];
      if (M->v_tendon_type[i] == OX_TEN_FIXED) {
        for (int a = 0; a < num; a++) J0[jd[M->v_wrap_objid[adr + a]]] += M->v_wrap_prm[adr + a];
      } else {
        double L0 = 0;
        std::vector<double> ja, jb, jr_;
        for (int a = 0; a + 1 < num; a++) {
          double p[2][3];
          int bodies[2];
          for (int s2 = 0; s2 < 2; s2++) {
            const int sid = M->v_wrap_objid[adr + a + s2];
            bodies[s2] = M->v_site_bodyid[sid];
            double r3[3];
            hm::rotvec(r3, &M->v_site_pos[3 * sid], &xquat[4 * bodies[s2]]);
            for (int c = 0; c < 3; c++) p[s2][c] = xpos[3 * bodies[s2] + c] + r3[c];
          }
          double dvec[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
          const double len = std::sqrt(dvec[0] * dvec[0] + dvec[1] * dvec[1] + dvec[2] * dvec[2]);
          L0 += len;
          if (len < OX_MINVAL) continue;
          jac(bodies[0], p[0], ja, jr_);
          jac(bodies[1], p[1], jb, jr_);
          for (int k = 0; k < nv; k++)
            for (int c = 0; c < 3; c++) J0[k] += dvec[c] / len * (jb[c * nv + k] - ja[c * nv + k]);
        }
        M->v_tendon_length0[i] = L0;
        if (spring_default[i]) { M->v_tendon_lengthspring[2 * i] = L0; M->v_tendon_lengthspring[2 * i + 1] = L0; }
      }
      double w = 0;
      for (int a = 0; a < nv; a++)
        for (int c = 0; c < nv; c++) w += J0[a] * Minv[a * nv + c] * J0[c];
      M->v_tendon_invweight0[i] = w;
    }
    for (int b = 1; b < nbody; b++) {
      if (M->v_body_weldid[b] == 0) continue;  // static
      jac(b, &xipos[3 * b], jp, jr);
      double tr_p = 0, tr_r = 0;
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < nv; i++)
          for (int j = 0; j < nv; j++) {
            tr_p += jp[c * nv + i] * Minv[i * nv + j] * jp[c * nv + j];
            tr_r += jr[c * nv + i] * Minv[i * nv + j] * jr[c * nv + j];
          }
      M->v_body_invweight0[2 * b] = tr_p / 3;
      M->v_body_invweight0[2 * b + 1] = tr_r / 3;
    }
  } else {
    t.meaninertia = 1;
  }

  // ---- mocap bodies (src/physics.rs:154-170): static children of the world whose pose comes from mjData.mocap_pos / mocap_quat
  M->v_body_mocapid.assign(nbody, -1);
  t.nmocap = 0;
  for (int i = 1; i < nbody; i++) {
    if (!B.bodies[i].mocap) continue;
    if (M->v_body_jntnum[i] > 0) cfail("body '" + B.bodies[i].name + "': a mocap body cannot have joints");
    M->v_body_mocapid[i] = t.nmocap++;
  }

  // ---- equality constraints (src/physics.rs:147-152 eq_active): connect (3 rows) and joint (1 row); weld / tendon / flex are refused
  t.neq = 0;
  {
    // body frames at qpos0 (joints contribute nothing at the reference configuration)
    std::vector<double> xpos(3 * nbody, 0), xquat(4 * nbody, 0);
    xquat[0] = 1;
    for (int i = 1; i < nbody; i++) {
      int p = M->v_body_parentid[i];
      double r[3];
      hm::rotvec(r, &M->v_body_pos[3 * i], &xquat[4 * p]);
      for (int k = 0; k < 3; k++) xpos[3 * i + k] = xpos[3 * p + k] + r[k];
      hm::mulquat(&xquat[4 * i], &xquat[4 * p], &M->v_body_quat[4 * i]);
      hm::normalize4(&xquat[4 * i]);
    }
    int eq_rows = 0;
    for (auto& ch : root->children)
      if (ch->name == "equality")
        for (auto& e : ch->children) {
          if (e->name != "connect" && e->name != "joint" && e->name != "weld") cfail("equality <" + e->name + "> is outside the supported subset (connect, weld, joint)");
          check_attrs(*e, "equality_common");
          Attrs a = merged(B.c, *e, "equality", "");
          const std::string ename = a.str_or("name", "");
          double solref[2] = {0.02, 1}, solimp[5] = {0.9, 0.95, 0.001, 0.5, 2}, data[11] = {0};
          a.vec("solref", solref, 2);
          a.vec("solimp", solimp, 5, true);
          int type, o1, o2 = -1;
          if (e->name == "connect") {
            type = OX_EQ_CONNECT;
            if (!a.has("body1") || !a.has("anchor")) pfail(*e, "connect requires body1 and anchor");
            o1 = find_name(nm[OX_OBJ_BODY], a.str("body1"));
            if (o1 < 0) cfail("equality '" + ename + "': unknown body '" + a.str("body1") + "'");
            o2 = 0;
            if (a.has("body2")) { o2 = find_name(nm[OX_OBJ_BODY], a.str("body2")); if (o2 < 0) cfail("equality '" + ename + "': unknown body '" + a.str("body2") + "'"); }
            a.vec("anchor", data, 3);
            // anchor in body2's frame such that both anchors coincide at qpos0
            double w[3], d2[3], q2c[4] = {xquat[4 * o2], -xquat[4 * o2 + 1], -xquat[4 * o2 + 2], -xquat[4 * o2 + 3]};
            hm::rotvec(w, data, &xquat[4 * o1]);
            for (int k = 0; k < 3; k++) d2[k] = xpos[3 * o1 + k] + w[k] - xpos[3 * o2 + k];
            hm::rotvec(data + 3, d2, q2c);
            eq_rows += 3;
          } else if (e->name == "weld") {
            // weld: body2 keeps a fixed pose relative to body1. data = [weld point in body1's frame, the same point in body2's frame
            // (MJCF anchor), relative orientation q_rel (target q2 = q1 * q_rel), torquescale]. relpose all zero / absent = the pose at qpos0.
            type = OX_EQ_WELD;
            if (!a.has("body1")) pfail(*e, "weld requires body1");
            o1 = find_name(nm[OX_OBJ_BODY], a.str("body1"));
            if (o1 < 0) cfail("equality '" + ename + "': unknown body '" + a.str("body1") + "'");
            o2 = 0;
            if (a.has("body2")) { o2 = find_name(nm[OX_OBJ_BODY], a.str("body2")); if (o2 < 0) cfail("equality '" + ename + "': unknown body '" + a.str("body2") + "'"); }
            double relpose[7] = {0, 0, 0, 0, 0, 0, 0}, anchor2[3] = {0, 0, 0}, qrel[4], rpos[3];
            a.vec("relpose", relpose, 7);
            a.vec("anchor", anchor2, 3);
            const double* q1 = &xquat[4 * o1];
            const double q1c[4] = {q1[0], -q1[1], -q1[2], -q1[3]};
            if (relpose[3] == 0 && relpose[4] == 0 && relpose[5] == 0 && relpose[6] == 0) {
              hm::mulquat(qrel, q1c, &xquat[4 * o2]);
              double dpos[3] = {xpos[3 * o2] - xpos[3 * o1], xpos[3 * o2 + 1] - xpos[3 * o1 + 1], xpos[3 * o2 + 2] - xpos[3 * o1 + 2]};
              hm::rotvec(rpos, dpos, q1c);
            } else {
              for (int k = 0; k < 4; k++) qrel[k] = relpose[3 + k];
              for (int k = 0; k < 3; k++) rpos[k] = relpose[k];
            }
            hm::normalize4(qrel);
            double w[3];
            hm::rotvec(w, anchor2, qrel);
            for (int k = 0; k < 3; k++) { data[k] = rpos[k] + w[k]; data[3 + k] = anchor2[k]; }
            for (int k = 0; k < 4; k++) data[6 + k] = qrel[k];
            data[10] = a.num("torquescale", 1.0);
            eq_rows += 6;
          } else {
            type = OX_EQ_JOINT;
            if (!a.has("joint1")) pfail(*e, "joint equality requires joint1");
            o1 = find_name(nm[OX_OBJ_JOINT], a.str("joint1"));
            if (o1 < 0) cfail("equality '" + ename + "': unknown joint '" + a.str("joint1") + "'");
            if (a.has("joint2")) { o2 = find_name(nm[OX_OBJ_JOINT], a.str("joint2")); if (o2 < 0) cfail("equality '" + ename + "': unknown joint '" + a.str("joint2") + "'"); }
            for (int j : {o1, o2})
              if (j >= 0 && B.joints[j].type != OX_JNT_HINGE && B.joints[j].type != OX_JNT_SLIDE) cfail("equality '" + ename + "': joint equalities couple hinge / slide joints");
            data[1] = 1;  // polycoef default "0 1 0 0 0"
            a.vec("polycoef", data, 5, true);
            eq_rows += 1;
          }
          const int act = a.boolean("active");
          M->v_eq_type.push_back(type); M->v_eq_obj1id.push_back(o1); M->v_eq_obj2id.push_back(o2); M->v_eq_active0.push_back(act == 0 ? 0 : 1);
          for (double v : solref) M->v_eq_solref.push_back(v);
          for (double v : solimp) M->v_eq_solimp.push_back(v);
          for (double v : data) M->v_eq_data.push_back(v);
          nm[OX_OBJ_EQUALITY].push_back(ename);
          t.neq++;
        }
    if (!(t.disableflags & (OX_DSBL_EQUALITY | OX_DSBL_CONSTRAINT))) t.nefcmax += eq_rows;
    check_unique(OX_OBJ_EQUALITY, "equality");
  }

  M->finalize();
  return M.release();
}

}  // namespace ox

void ox_model::finalize() {
#define OX_X(name, n, w)                                                                              \
  if ((long)v_##name.size() != (long)t.n * (w))                                                       \
    throw ox::CompileError(std::string("internal: table ") + #name + " has wrong length");            \
  t.name = v_##name.data();
  OX_MODEL_INT_TABLES(OX_X)
  OX_MODEL_REAL_TABLES(OX_X)
#undef OX_X
}
