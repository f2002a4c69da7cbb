// rbd_urdf.cpp — URDF text -> rbd_model_desc, for batch use without pinocchio.
//
// Stands for what the reference's loader obtains from `pinocchio::urdf::buildModel(urdfFilename, [JointModelFreeFlyer()],
// model)` (URDFModelLoader.cpp:126-139) followed by its two frame tables: every frame that exists after parsing becomes
// an "extra frame" (URDFModelLoader.cpp:159-165) and one OP_FRAME "Body_<j>_CoM" per joint, universe included, is
// appended at the joint's CoM lever with identity rotation (URDFModelLoader.cpp:175-183).  The mapped outputs are the
// CoM frames first, then the extra frames (KinematicChainMapping.inl:382-395).
//
// pinocchio / urdfdom are not available in this image; the conventions below restate pinocchio 3.3.1's URDF visitor
// (parsers/urdf/model.hxx) and urdfdom 4.0.1's tree construction from memory and are labelled so in DESIGN.md:
//   * urdfdom keeps joints in a std::map<name, joint>, so a link's children are visited in joint-name order;
//   * the root link's inertia goes to the universe (fixed base: frames "root_joint" FIXED_JOINT + body frame) or to a
//     free-flyer "root_joint" (JOINT frame + body frame);
//   * revolute / continuous / prismatic joints: a joint at parentFramePlacement * origin, a JOINT frame and a BODY frame
//     at identity; axis == e_x / e_y / e_z (Eigen isApprox, 1e-12) picks the aligned joint model, anything else the
//     unaligned one with the normalised axis; continuous joints are the unbounded (cos, sin) models;
//   * fixed joints add no joint: a FIXED_JOINT frame + BODY frame at parentFramePlacement * origin on the parent's joint,
//     the child's inertia is appended to that joint (Model::appendBodyToJoint: inertias[j] += placement.act(Y));
//   * <mimic>, <limit>, <dynamics>, <visual>, <collision> do not reach the hot path and are ignored (pinocchio 3.3.1
//     does not model mimic joints); "planar" joints are rejected with RBD_ERR_UNSUPPORTED.
// The XML subset parser below handles declarations, comments, DOCTYPE, CDATA, both quote styles and the five predefined
// entities — everything the reference's own URDF files (and xacro output in general) contain.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <charconv>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../../include/rbd_b200.h"

int rbd_internal_set_error(int code, const char* msg);  // rbd_capi.cu (thread-local error string)

namespace {

// ------------------------------------------------------------------------------------------------ XML subset
struct XmlNode {
  std::string tag;
  std::vector<std::pair<std::string, std::string>> attrs;
  std::vector<std::unique_ptr<XmlNode>> kids;
  const std::string* attr(const char* name) const {
    for (const auto& a : attrs)
      if (a.first == name) return &a.second;
    return nullptr;
  }
  const XmlNode* child(const char* t) const {
    for (const auto& k : kids)
      if (k->tag == t) return k.get();
    return nullptr;
  }
};

struct XmlError {
  std::string msg;
};

class XmlParser {
 public:
  explicit XmlParser(const char* s) : p_(s), begin_(s) {}
  std::unique_ptr<XmlNode> parse_document() {
    skip_misc();
    if (*p_ != '<') fail("no root element");
    std::unique_ptr<XmlNode> root = parse_element(0);
    skip_misc();
    if (*p_) fail("content after the root element");
    return root;
  }

 private:
  const char* p_;
  const char* begin_;
  [[noreturn]] void fail(const std::string& m) {
    int line = 1;
    for (const char* c = begin_; c < p_; ++c) line += *c == '\n';
    throw XmlError{"XML: " + m + " (line " + std::to_string(line) + ")"};
  }
  static bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
  static bool is_name(char c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == ':' || c == '-' ||
           c == '.' || (static_cast<unsigned char>(c) & 0x80);
  }
  void skip_space() {
    while (is_space(*p_)) ++p_;
  }
  bool starts(const char* s) const { return strncmp(p_, s, strlen(s)) == 0; }
  void skip_until(const char* end, const char* what) {
    const char* e = strstr(p_, end);
    if (!e) fail(std::string("unterminated ") + what);
    p_ = e + strlen(end);
  }
  // whitespace, byte-order mark, <?...?>, <!--...-->, <!DOCTYPE ...> between elements
  void skip_misc() {
    if (p_ == begin_ && starts("\xEF\xBB\xBF")) p_ += 3;
    for (;;) {
      skip_space();
      if (starts("<?")) skip_until("?>", "processing instruction");
      else if (starts("<!--")) skip_until("-->", "comment");
      else if (starts("<!DOCTYPE")) {
        int depth = 0;
        for (; *p_; ++p_) {
          if (*p_ == '[') ++depth;
          else if (*p_ == ']') --depth;
          else if (*p_ == '>' && depth <= 0) break;
        }
        if (!*p_) fail("unterminated DOCTYPE");
        ++p_;
      } else return;
    }
  }
  std::string parse_name() {
    const char* s = p_;
    while (is_name(*p_)) ++p_;
    if (p_ == s) fail("name expected");
    return std::string(s, p_);
  }
  std::string parse_value() {
    const char q = *p_;
    if (q != '"' && q != '\'') fail("quoted attribute value expected");
    ++p_;
    std::string v;
    while (*p_ && *p_ != q) {
      if (*p_ == '&') {
        static const struct { const char* ent; char c; } ents[] = {
            {"&lt;", '<'}, {"&gt;", '>'}, {"&amp;", '&'}, {"&quot;", '"'}, {"&apos;", '\''}};
        bool hit = false;
        for (const auto& e : ents)
          if (starts(e.ent)) { v.push_back(e.c); p_ += strlen(e.ent); hit = true; break; }
        if (!hit) {
          if (starts("&#")) {  // numeric character reference: ASCII only (URDF attribute values are names and numbers)
            char* end = nullptr;
            const long code = p_[2] == 'x' ? strtol(p_ + 3, &end, 16) : strtol(p_ + 2, &end, 10);
            if (!end || *end != ';' || code <= 0 || code > 127) fail("unsupported character reference");
            v.push_back(static_cast<char>(code));
            p_ = end + 1;
          } else fail("unknown entity");
        }
      } else v.push_back(*p_++);
    }
    if (*p_ != q) fail("unterminated attribute value");
    ++p_;
    return v;
  }
  std::unique_ptr<XmlNode> parse_element(int depth) {
    if (depth > 256) fail("elements nested too deeply");
    ++p_;  // '<'
    std::unique_ptr<XmlNode> n(new XmlNode);
    n->tag = parse_name();
    for (;;) {
      skip_space();
      if (*p_ == '/') {
        if (p_[1] != '>') fail("'/>' expected");
        p_ += 2;
        return n;
      }
      if (*p_ == '>') { ++p_; break; }
      if (!*p_) fail("unterminated start tag");
      std::string an = parse_name();
      skip_space();
      if (*p_ != '=') fail("'=' expected after attribute name");
      ++p_;
      skip_space();
      n->attrs.emplace_back(std::move(an), parse_value());
    }
    for (;;) {  // content: child elements; text is of no interest to URDF
      while (*p_ && *p_ != '<') ++p_;
      if (!*p_) fail("unterminated element <" + n->tag + ">");
      if (starts("<!--")) skip_until("-->", "comment");
      else if (starts("<![CDATA[")) skip_until("]]>", "CDATA section");
      else if (starts("<?")) skip_until("?>", "processing instruction");
      else if (p_[1] == '/') {
        p_ += 2;
        const std::string t = parse_name();
        if (t != n->tag) fail("</" + t + "> closes <" + n->tag + ">");
        skip_space();
        if (*p_ != '>') fail("'>' expected");
        ++p_;
        return n;
      } else n->kids.push_back(parse_element(depth + 1));
    }
  }
};

// ------------------------------------------------------------------------------------------------ spatial helpers
// SE3: 12 doubles, R row-major (9) then p (3); inertia: [m, cx, cy, cz, Ixx, Ixy, Iyy, Ixz, Iyz, Izz]
struct SE3 {
  double v[12];
};
struct Inertia {
  double v[10];
};
SE3 se3_identity() { return SE3{{1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0}}; }
Inertia inertia_zero() { return Inertia{{0, 0, 0, 0, 0, 0, 0, 0, 0, 0}}; }
void mat3_mul(const double* a, const double* b, double* c) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(c, t, sizeof t);
}
SE3 se3_mul(const SE3& a, const SE3& b) {
  SE3 r;
  mat3_mul(a.v, b.v, r.v);
  for (int i = 0; i < 3; ++i)
    r.v[9 + i] = a.v[3 * i] * b.v[9] + a.v[3 * i + 1] * b.v[10] + a.v[3 * i + 2] * b.v[11] + a.v[9 + i];
  return r;
}
// URDF fixed-axis roll / pitch / yaw: R = Rz(y) Ry(p) Rx(r)
void rpy_to_R(double r, double p, double y, double* R) {
  const double cr = cos(r), sr = sin(r), cp = cos(p), sp = sin(p), cy = cos(y), sy = sin(y);
  const double Rx[9] = {1, 0, 0, 0, cr, -sr, 0, sr, cr};
  const double Ry[9] = {cp, 0, sp, 0, 1, 0, -sp, 0, cp};
  const double Rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
  double t[9];
  mat3_mul(Rz, Ry, t);
  mat3_mul(t, Rx, R);
}
void inertia_unpack(const Inertia& y, double& m, double* c, double* I) {
  m = y.v[0];
  c[0] = y.v[1]; c[1] = y.v[2]; c[2] = y.v[3];
  I[0] = y.v[4]; I[1] = y.v[5]; I[2] = y.v[7];
  I[3] = y.v[5]; I[4] = y.v[6]; I[5] = y.v[8];
  I[6] = y.v[7]; I[7] = y.v[8]; I[8] = y.v[9];
}
Inertia inertia_pack(double m, const double* c, const double* I) {
  return Inertia{{m, c[0], c[1], c[2], I[0], I[1], I[4], I[2], I[5], I[8]}};
}
// pinocchio InertiaTpl::__plus__ (mass floor at machine epsilon)
Inertia inertia_add(const Inertia& ya, const Inertia& yb) {
  double ma, mb, ca[3], cb[3], Ia[9], Ib[9];
  inertia_unpack(ya, ma, ca, Ia);
  inertia_unpack(yb, mb, cb, Ib);
  const double mab = ma + mb;
  const double inv = 1.0 / std::max(mab, 2.220446049250313e-16);
  const double ab[3] = {ca[0] - cb[0], ca[1] - cb[1], ca[2] - cb[2]};
  const double S[9] = {0, -ab[2], ab[1], ab[2], 0, -ab[0], -ab[1], ab[0], 0};
  double SS[9], c[3], I[9];
  mat3_mul(S, S, SS);
  const double f = ma * mb * inv;
  for (int k = 0; k < 9; ++k) I[k] = Ia[k] + Ib[k] - f * SS[k];
  for (int k = 0; k < 3; ++k) c[k] = (ma * ca[k] + mb * cb[k]) * inv;
  return inertia_pack(mab, c, I);
}
// pinocchio InertiaTpl::se3Action_impl: inertia given in frame b expressed in frame a (M = aMb)
Inertia inertia_act(const SE3& M, const Inertia& y) {
  double m, c[3], I[9], Rt[9], t[9], RIRt[9], c2[3];
  inertia_unpack(y, m, c, I);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Rt[3 * i + j] = M.v[3 * j + i];
  mat3_mul(M.v, I, t);
  mat3_mul(t, Rt, RIRt);
  for (int i = 0; i < 3; ++i) c2[i] = M.v[9 + i] + (M.v[3 * i] * c[0] + M.v[3 * i + 1] * c[1] + M.v[3 * i + 2] * c[2]);
  return inertia_pack(m, c2, RIRt);
}

// ------------------------------------------------------------------------------------------------ builder
enum { F_OP = 1, F_JOINT = 2, F_FIXED_JOINT = 4, F_BODY = 8 };  // pinocchio::FrameType bit values

const int kJointNq[14] = {0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 7};
const int kJointNv[14] = {0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 6};

struct UrdfError {
  int code;
  std::string msg;
};

}  // namespace

struct rbd_urdf_model {
  std::string name;
  std::vector<int32_t> parents, jtype, idx_q, idx_v, frame_parent, frame_type, out_frames;
  std::vector<double> axis, placement, inertia, frame_placement;
  std::vector<std::string> joint_names, frame_names;
  int32_t nframes0 = 0;
  rbd_model_desc desc{};

  rbd_urdf_model() {
    // Model(): joint 0 "universe" and frame 0 "universe" (FIXED_JOINT on joint 0)
    add_joint_raw(0, RBD_JOINT_UNIVERSE, se3_identity(), "universe", nullptr);
    add_frame("universe", 0, se3_identity(), F_FIXED_JOINT);
  }
  int add_joint_raw(int parent, int jt, const SE3& pl, const std::string& nm, const double* ax) {
    static const double unit[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double a[3] = {0, 0, 0};
    if (ax) memcpy(a, ax, sizeof a);
    else if (jt >= RBD_JOINT_RX && jt <= RBD_JOINT_RUBZ && jt != RBD_JOINT_RU && jt != RBD_JOINT_PU) {
      const int k = jt <= RBD_JOINT_RZ ? jt - RBD_JOINT_RX : jt <= RBD_JOINT_PZ ? jt - RBD_JOINT_PX : jt - RBD_JOINT_RUBX;
      memcpy(a, unit[k], sizeof a);
    }
    parents.push_back(parent);
    jtype.push_back(jt);
    axis.insert(axis.end(), a, a + 3);
    placement.insert(placement.end(), pl.v, pl.v + 12);
    const Inertia z = inertia_zero();
    inertia.insert(inertia.end(), z.v, z.v + 10);
    joint_names.push_back(nm);
    return static_cast<int>(parents.size()) - 1;
  }
  int add_frame(const std::string& nm, int parent_joint, const SE3& pl, int type) {
    frame_parent.push_back(parent_joint);
    frame_placement.insert(frame_placement.end(), pl.v, pl.v + 12);
    frame_names.push_back(nm);
    frame_type.push_back(type);
    return static_cast<int>(frame_parent.size()) - 1;
  }
  SE3 frame_pl(int f) const {
    SE3 s;
    memcpy(s.v, &frame_placement[12 * f], sizeof s.v);
    return s;
  }
  void append_body(int joint, const Inertia& y, const SE3& pl) {
    Inertia cur;
    memcpy(cur.v, &inertia[10 * joint], sizeof cur.v);
    const Inertia sum = inertia_add(cur, inertia_act(pl, y));
    memcpy(&inertia[10 * joint], sum.v, sizeof sum.v);
  }
  void finish() {
    const int nj = static_cast<int>(parents.size());
    idx_q.assign(nj, 0);
    idx_v.assign(nj, 0);
    int iq = 0, iv = 0;
    for (int i = 1; i < nj; ++i) {
      idx_q[i] = iq; idx_v[i] = iv;
      iq += kJointNq[jtype[i]]; iv += kJointNv[jtype[i]];
    }
    nframes0 = static_cast<int32_t>(frame_parent.size());
    std::vector<int32_t> com;
    for (int j = 0; j < nj; ++j) {  // URDFModelLoader.cpp:175-183
      SE3 pl = se3_identity();
      for (int k = 0; k < 3; ++k) pl.v[9 + k] = inertia[10 * j + 1 + k];
      com.push_back(add_frame("Body_" + std::to_string(j) + "_CoM", j, pl, F_OP));
    }
    out_frames = com;
    for (int f = 0; f < nframes0; ++f) out_frames.push_back(f);  // URDFModelLoader.cpp:159-165
    desc.njoints = nj; desc.nq = iq; desc.nv = iv;
    desc.parents = parents.data(); desc.jtype = jtype.data(); desc.axis = axis.data();
    desc.placement = placement.data(); desc.idx_q = idx_q.data(); desc.idx_v = idx_v.data();
    desc.inertia = inertia.data();
    desc.nframes = static_cast<int32_t>(frame_parent.size());
    desc.frame_parent = frame_parent.data(); desc.frame_placement = frame_placement.data();
    desc.gravity[0] = 0.0; desc.gravity[1] = 0.0; desc.gravity[2] = -9.81;  // pinocchio::Model::gravity981
    desc.n_out = static_cast<int32_t>(out_frames.size());
    desc.out_frames = out_frames.data();
  }
};

namespace {

// Numbers are parsed with std::from_chars, not strtod: a host application (SOFA's GUI) may have set a LC_NUMERIC locale
// with a decimal comma, under which strtod("0.5") stops at the dot.
void parse_floats(const std::string* s, int n, double* out, const char* what) {
  for (int k = 0; k < n; ++k) out[k] = 0.0;
  if (!s) return;
  const char* p = s->c_str();
  const char* const last = p + s->size();
  for (int k = 0; k < n; ++k) {
    while (p < last && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    if (p < last && *p == '+') ++p;  // from_chars takes no leading plus sign
    const std::from_chars_result r = std::from_chars(p, last, out[k]);
    if (r.ec != std::errc() || r.ptr == p)
      throw UrdfError{RBD_ERR_MODEL, std::string("URDF: ") + what + " needs " + std::to_string(n) + " numbers, got \"" + *s + "\""};
    p = r.ptr;
  }
  while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') ++p;
  if (*p) throw UrdfError{RBD_ERR_MODEL, std::string("URDF: ") + what + " has trailing text: \"" + *s + "\""};
}
double parse_float(const std::string* s, const char* what) {
  double v = 0.0;
  parse_floats(s, 1, &v, what);
  return v;
}
SE3 parse_origin(const XmlNode* el) {
  SE3 s = se3_identity();
  if (!el) return s;
  double xyz[3], rpy[3];
  parse_floats(el->attr("xyz"), 3, xyz, "origin xyz");
  parse_floats(el->attr("rpy"), 3, rpy, "origin rpy");
  rpy_to_R(rpy[0], rpy[1], rpy[2], s.v);
  memcpy(s.v + 9, xyz, sizeof xyz);
  return s;
}
// URDF <inertial> -> Inertia(mass, com, R I R^T) (pinocchio convertFromUrdf); a link without one weighs nothing
Inertia link_inertia(const XmlNode* link) {
  const XmlNode* ine = link->child("inertial");
  if (!ine) return inertia_zero();
  const SE3 o = parse_origin(ine->child("origin"));
  const XmlNode* me = ine->child("mass");
  const XmlNode* ie = ine->child("inertia");
  if (!me || !me->attr("value")) throw UrdfError{RBD_ERR_MODEL, "URDF: <inertial> without <mass value=...>"};
  const double m = parse_float(me->attr("value"), "mass value");
  double g[6] = {0, 0, 0, 0, 0, 0};
  if (ie) {
    static const char* keys[6] = {"ixx", "ixy", "ixz", "iyy", "iyz", "izz"};
    for (int k = 0; k < 6; ++k)
      if (ie->attr(keys[k])) g[k] = parse_float(ie->attr(keys[k]), keys[k]);
  }
  const double I[9] = {g[0], g[1], g[2], g[1], g[3], g[4], g[2], g[4], g[5]};
  double Rt[9], t[9], RIRt[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Rt[3 * i + j] = o.v[3 * j + i];
  mat3_mul(o.v, I, t);
  mat3_mul(t, Rt, RIRt);
  return inertia_pack(m, o.v + 9, RIRt);
}
// Eigen isApprox(unit vector), default precision 1e-12: |a - e|^2 <= prec^2 * min(|a|^2, 1)
bool approx_unit(const double* a, int k) {
  double d2 = 0, n2 = 0;
  for (int i = 0; i < 3; ++i) {
    const double e = i == k ? 1.0 : 0.0;
    d2 += (a[i] - e) * (a[i] - e);
    n2 += a[i] * a[i];
  }
  return d2 <= 1e-24 * std::min(n2, 1.0);
}

struct UrdfJoint {
  const XmlNode* el;
  std::string parent, child;
};

void build_from_xml(const XmlNode& robot, bool free_flyer, rbd_urdf_model& b) {
  if (robot.tag != "robot") throw UrdfError{RBD_ERR_MODEL, "URDF: root element is <" + robot.tag + ">, not <robot>"};
  if (robot.attr("name")) b.name = *robot.attr("name");
  std::map<std::string, const XmlNode*> links;
  std::map<std::string, UrdfJoint> joints;  // sorted by name, as urdfdom's ModelInterface::joints_
  for (const auto& k : robot.kids) {
    if (k->tag == "link") {
      if (!k->attr("name")) throw UrdfError{RBD_ERR_MODEL, "URDF: <link> without a name"};
      if (!links.emplace(*k->attr("name"), k.get()).second)
        throw UrdfError{RBD_ERR_MODEL, "URDF: link \"" + *k->attr("name") + "\" is defined twice"};
    } else if (k->tag == "joint") {
      if (!k->attr("name")) throw UrdfError{RBD_ERR_MODEL, "URDF: <joint> without a name"};
      const XmlNode* pe = k->child("parent");
      const XmlNode* ce = k->child("child");
      if (!pe || !ce || !pe->attr("link") || !ce->attr("link"))
        throw UrdfError{RBD_ERR_MODEL, "URDF: joint \"" + *k->attr("name") + "\" lacks <parent link> / <child link>"};
      if (!joints.emplace(*k->attr("name"), UrdfJoint{k.get(), *pe->attr("link"), *ce->attr("link")}).second)
        throw UrdfError{RBD_ERR_MODEL, "URDF: joint \"" + *k->attr("name") + "\" is defined twice"};
    }
  }
  if (links.empty()) throw UrdfError{RBD_ERR_MODEL, "URDF: no links"};
  std::map<std::string, std::vector<const std::string*>> child_joints;  // parent link -> joint names (sorted)
  std::map<std::string, int> has_parent;
  for (const auto& kv : joints) {
    const UrdfJoint& j = kv.second;
    if (!links.count(j.parent) || !links.count(j.child))
      throw UrdfError{RBD_ERR_MODEL, "URDF: joint \"" + kv.first + "\" names an unknown link"};
    if (has_parent[j.child]++) throw UrdfError{RBD_ERR_MODEL, "URDF: link \"" + j.child + "\" has two parent joints"};
    child_joints[j.parent].push_back(&kv.first);
  }
  std::string root_link;
  for (const auto& kv : links)
    if (!has_parent.count(kv.first)) {
      if (!root_link.empty()) throw UrdfError{RBD_ERR_MODEL, "URDF: two root links (\"" + root_link + "\", \"" + kv.first + "\")"};
      root_link = kv.first;
    }
  if (root_link.empty()) throw UrdfError{RBD_ERR_MODEL, "URDF: no root link (kinematic loop)"};

  std::map<std::string, int> body_frame;
  const Inertia Yroot = link_inertia(links[root_link]);
  if (free_flyer) {  // UrdfVisitorWithRootJoint::addRootJoint
    const int jid = b.add_joint_raw(0, RBD_JOINT_FF, se3_identity(), "root_joint", nullptr);
    b.add_frame("root_joint", jid, se3_identity(), F_JOINT);
    b.append_body(jid, Yroot, se3_identity());
    body_frame[root_link] = b.add_frame(root_link, jid, se3_identity(), F_BODY);
  } else {  // UrdfVisitor::addRootJoint -> addFixedJointAndBody(0, I, "root_joint", Y, root)
    b.add_frame("root_joint", 0, se3_identity(), F_FIXED_JOINT);
    b.append_body(0, Yroot, se3_identity());
    body_frame[root_link] = b.add_frame(root_link, 0, se3_identity(), F_BODY);
  }

  // depth-first over child links, explicit stack (URDF chains can be long)
  struct Item { const std::string* link; size_t next; };
  std::vector<Item> stack;
  stack.push_back(Item{&root_link, 0});
  size_t visited = 1;
  while (!stack.empty()) {
    Item& it = stack.back();
    const auto cj = child_joints.find(*it.link);
    if (cj == child_joints.end() || it.next >= cj->second.size()) { stack.pop_back(); continue; }
    const std::string& jn = *cj->second[it.next++];
    const UrdfJoint& j = joints[jn];
    const std::string* kind = j.el->attr("type");
    if (!kind) throw UrdfError{RBD_ERR_MODEL, "URDF: joint \"" + jn + "\" has no type"};
    const SE3 pl = parse_origin(j.el->child("origin"));
    const Inertia Yc = link_inertia(links[j.child]);
    const int pf = body_frame[j.parent];
    const int pj = b.frame_parent[pf];
    const SE3 jpl = se3_mul(b.frame_pl(pf), pl);
    if (*kind == "revolute" || *kind == "continuous" || *kind == "prismatic") {
      double ax[3] = {1, 0, 0};  // URDF default axis
      const XmlNode* ae = j.el->child("axis");
      if (ae && ae->attr("xyz")) parse_floats(ae->attr("xyz"), 3, ax, "axis xyz");
      const int base = *kind == "revolute" ? RBD_JOINT_RX : *kind == "continuous" ? RBD_JOINT_RUBX : RBD_JOINT_PX;
      int jt = base + 3;
      for (int k = 0; k < 3; ++k)
        if (approx_unit(ax, k)) jt = base + k;
      double unit[3] = {0, 0, 0};
      if (jt == base + 3) {
        const double nrm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        if (!(nrm > 0.0)) throw UrdfError{RBD_ERR_MODEL, "URDF: joint \"" + jn + "\" has a zero axis"};
        for (int k = 0; k < 3; ++k) unit[k] = ax[k] / nrm;
      } else unit[jt - base] = 1.0;
      const int jid = b.add_joint_raw(pj, jt, jpl, jn, unit);
      b.add_frame(jn, jid, se3_identity(), F_JOINT);
      b.append_body(jid, Yc, se3_identity());
      body_frame[j.child] = b.add_frame(j.child, jid, se3_identity(), F_BODY);
    } else if (*kind == "floating") {
      const int jid = b.add_joint_raw(pj, RBD_JOINT_FF, jpl, jn, nullptr);
      b.add_frame(jn, jid, se3_identity(), F_JOINT);
      b.append_body(jid, Yc, se3_identity());
      body_frame[j.child] = b.add_frame(j.child, jid, se3_identity(), F_BODY);
    } else if (*kind == "fixed") {
      b.add_frame(jn, pj, jpl, F_FIXED_JOINT);
      b.append_body(pj, Yc, jpl);
      body_frame[j.child] = b.add_frame(j.child, pj, jpl, F_BODY);
    } else if (*kind == "planar") {
      throw UrdfError{RBD_ERR_UNSUPPORTED, "URDF: joint \"" + jn + "\" is planar; the back end has no planar joint model"};
    } else {
      throw UrdfError{RBD_ERR_MODEL, "URDF: joint \"" + jn + "\" has unknown type \"" + *kind + "\""};
    }
    ++visited;
    stack.push_back(Item{&j.child, 0});
  }
  if (visited != links.size()) throw UrdfError{RBD_ERR_MODEL, "URDF: some links are not reachable from the root link"};
  b.finish();
}

int load_text(const char* text, int free_flyer, rbd_urdf_model** out) {
  if (!text || !out) return rbd_internal_set_error(RBD_ERR_INVALID_ARG, "rbd_urdf_load: null argument");
  *out = nullptr;
  try {
    XmlParser xp(text);
    std::unique_ptr<XmlNode> root = xp.parse_document();
    std::unique_ptr<rbd_urdf_model> m(new rbd_urdf_model);
    build_from_xml(*root, free_flyer != 0, *m);
    *out = m.release();
    return RBD_OK;
  } catch (const XmlError& e) {
    return rbd_internal_set_error(RBD_ERR_MODEL, e.msg.c_str());
  } catch (const UrdfError& e) {
    return rbd_internal_set_error(e.code, e.msg.c_str());
  } catch (const std::bad_alloc&) {
    return rbd_internal_set_error(RBD_ERR_ALLOC, "rbd_urdf_load: out of memory");
  } catch (...) {
    return rbd_internal_set_error(RBD_ERR_MODEL, "rbd_urdf_load: unexpected failure");
  }
}

}  // namespace

extern "C" {

RBD_API int rbd_urdf_load_string(const char* xml, int free_flyer, rbd_urdf_model** out) { return load_text(xml, free_flyer, out); }

RBD_API int rbd_urdf_load_file(const char* path, int free_flyer, rbd_urdf_model** out) {
  if (!path || !out) return rbd_internal_set_error(RBD_ERR_INVALID_ARG, "rbd_urdf_load_file: null argument");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return rbd_internal_set_error(RBD_ERR_INVALID_ARG, (std::string("rbd_urdf_load_file: cannot open ") + path).c_str());
  std::string text;
  try {
    char buf[65536];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
  } catch (...) {
    fclose(f);
    return rbd_internal_set_error(RBD_ERR_ALLOC, "rbd_urdf_load_file: out of memory");
  }
  fclose(f);
  if (text.find('\0') != std::string::npos) return rbd_internal_set_error(RBD_ERR_MODEL, "rbd_urdf_load_file: not a text file");
  return load_text(text.c_str(), free_flyer, out);
}

RBD_API void rbd_urdf_destroy(rbd_urdf_model* m) { delete m; }

RBD_API const rbd_model_desc* rbd_urdf_desc(const rbd_urdf_model* m) { return m ? &m->desc : nullptr; }

RBD_API int32_t rbd_urdf_nframes0(const rbd_urdf_model* m) { return m ? m->nframes0 : -1; }

RBD_API const char* rbd_urdf_robot_name(const rbd_urdf_model* m) { return m ? m->name.c_str() : nullptr; }

RBD_API const char* rbd_urdf_joint_name(const rbd_urdf_model* m, int32_t joint) {
  return (m && joint >= 0 && joint < static_cast<int32_t>(m->joint_names.size())) ? m->joint_names[joint].c_str() : nullptr;
}

RBD_API const char* rbd_urdf_frame_name(const rbd_urdf_model* m, int32_t frame) {
  return (m && frame >= 0 && frame < static_cast<int32_t>(m->frame_names.size())) ? m->frame_names[frame].c_str() : nullptr;
}

RBD_API int32_t rbd_urdf_frame_type(const rbd_urdf_model* m, int32_t frame) {
  return (m && frame >= 0 && frame < static_cast<int32_t>(m->frame_type.size())) ? m->frame_type[frame] : -1;
}

// index of the first frame called `name` whose type matches `type_mask` (pinocchio Model::getFrameId), or -1
RBD_API int32_t rbd_urdf_find_frame(const rbd_urdf_model* m, const char* name, int32_t type_mask) {
  if (!m || !name) return -1;
  for (size_t f = 0; f < m->frame_names.size(); ++f)
    if ((m->frame_type[f] & type_mask) && m->frame_names[f] == name) return static_cast<int32_t>(f);
  return -1;
}

RBD_API int32_t rbd_urdf_find_joint(const rbd_urdf_model* m, const char* name) {
  if (!m || !name) return -1;
  for (size_t j = 0; j < m->joint_names.size(); ++j)
    if (m->joint_names[j] == name) return static_cast<int32_t>(j);
  return -1;
}

}  // extern "C"
