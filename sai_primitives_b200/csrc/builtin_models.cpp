// Built-in robot models of the product, written independently of oracle/robots.py
// (tests/test_models.py cross-checks the two).  A model is assembled the way RBDL's URDF
// reader hands it to sai-model: links attached through fixed joints are merged into their
// parent body.
//
// Physical parameters are the ones published in the reference's URDF data files:
//   panda               examples/15-haptic_control_impedance_type/panda_arm.urdf:4-183
//   panda_sliding_base  examples/06-partial_joint_task/panda_arm_sliding_base.urdf:171-233
//   rrrr                examples/11-planar_robot_controller/rrrrbot.urdf:5-166
//   puma_like           authored here (the reference's puma.urdf lives in sai-model)
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sai_b200_osc.h"

namespace {

struct V3 {
	double x, y, z;
};
struct M3 {
	double m[9];
};

M3 ident() { return M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
M3 mul(const M3& a, const M3& b) {
	M3 c{};
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) {
			double s = 0;
			for (int k = 0; k < 3; k++) s += a.m[3 * i + k] * b.m[3 * k + j];
			c.m[3 * i + j] = s;
		}
	return c;
}
M3 transpose(const M3& a) {
	M3 c{};
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) c.m[3 * i + j] = a.m[3 * j + i];
	return c;
}
V3 mul(const M3& a, const V3& v) {
	return V3{a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
			  a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
V3 add(const V3& a, const V3& b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
// URDF rpy: R = Rz(yaw) Ry(pitch) Rx(roll)
M3 rpy(double r, double p, double y) {
	const double cr = std::cos(r), sr = std::sin(r), cp = std::cos(p), sp = std::sin(p), cy = std::cos(y), sy = std::sin(y);
	return M3{{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, sy * cp, sy * sp * sr + cy * cr,
			   sy * sp * cr - cy * sr, -sp, cp * sr, cp * cr}};
}

enum JType { FIXED = -1, REVOLUTE = 0, PRISMATIC = 1 };

struct Link {
	const char* name;
	JType jtype;
	V3 xyz;
	V3 rpy_;
	V3 axis;
	double lower, upper, velocity, effort;
	double mass;
	V3 com;
	V3 inertia_diag;
	V3 inertia_off{0, 0, 0};  // ixy, ixz, iyz (URDF loader; the built-in tables are diagonal)
	V3 com_rpy{0, 0, 0};	  // orientation of the inertial frame in the link frame
};

struct Built {
	osc_model_desc desc;
	std::map<std::string, osc_link_frame> frames;
};

// second moment about the origin of a body with inertia Ic (about its com c), mass m
void add_second_moment(double acc[9], const M3& Ic, double m, const V3& c) {
	const double cc = c.x * c.x + c.y * c.y + c.z * c.z;
	const double cv[3] = {c.x, c.y, c.z};
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) acc[3 * i + j] += Ic.m[3 * i + j] + m * ((i == j ? cc : 0.0) - cv[i] * cv[j]);
}

Built build(const std::vector<Link>& links) {
	Built b;
	std::memset(&b.desc, 0, sizeof(b.desc));
	osc_model_desc& d = b.desc;
	const M3 I = ident();
	std::memcpy(d.R_world_base, I.m, sizeof(I.m));
	d.gravity_world[2] = -9.81;
	int body = -1;
	M3 R_lb = ident();	// current link frame in the current body frame
	V3 t_lb{0, 0, 0};
	double am = 0;		// accumulators of the current body, about the body origin
	V3 ah{0, 0, 0};
	double aI[9] = {0};
	auto finish = [&](int bi) {
		const V3 c{ah.x / am, ah.y / am, ah.z / am};
		d.mass[bi] = am;
		d.com[bi][0] = c.x;
		d.com[bi][1] = c.y;
		d.com[bi][2] = c.z;
		const double cc = c.x * c.x + c.y * c.y + c.z * c.z;
		const double cv[3] = {c.x, c.y, c.z};
		for (int i = 0; i < 3; i++)
			for (int j = 0; j < 3; j++) d.inertia[bi][3 * i + j] = aI[3 * i + j] - am * ((i == j ? cc : 0.0) - cv[i] * cv[j]);
	};
	for (const Link& l : links) {
		const M3 Rj = rpy(l.rpy_.x, l.rpy_.y, l.rpy_.z);
		if (l.jtype == FIXED) {
			t_lb = add(t_lb, mul(R_lb, l.xyz));
			R_lb = mul(R_lb, Rj);
		} else {
			if (body >= 0) finish(body);
			body++;
			d.jtype[body] = (int)l.jtype;
			const double an = std::sqrt(l.axis.x * l.axis.x + l.axis.y * l.axis.y + l.axis.z * l.axis.z);
			d.axis[body][0] = l.axis.x / an;
			d.axis[body][1] = l.axis.y / an;
			d.axis[body][2] = l.axis.z / an;
			const V3 tf = add(t_lb, mul(R_lb, l.xyz));
			const M3 Rf = mul(R_lb, Rj);
			std::memcpy(d.R_fix[body], Rf.m, sizeof(Rf.m));
			d.t_fix[body][0] = tf.x;
			d.t_fix[body][1] = tf.y;
			d.t_fix[body][2] = tf.z;
			d.q_lower[body] = l.lower;
			d.q_upper[body] = l.upper;
			d.dq_max[body] = l.velocity;
			d.effort[body] = l.effort;
			R_lb = ident();
			t_lb = V3{0, 0, 0};
			am = 0;
			ah = V3{0, 0, 0};
			std::memset(aI, 0, sizeof(aI));
		}
		osc_link_frame f;
		f.body = body;
		std::memcpy(f.R, R_lb.m, sizeof(R_lb.m));
		f.t[0] = t_lb.x;
		f.t[1] = t_lb.y;
		f.t[2] = t_lb.z;
		b.frames[l.name] = f;
		if (body >= 0) {
			const V3 c = add(t_lb, mul(R_lb, l.com));
			M3 Id{{l.inertia_diag.x, l.inertia_off.x, l.inertia_off.y, l.inertia_off.x, l.inertia_diag.y, l.inertia_off.z, l.inertia_off.y,
				   l.inertia_off.z, l.inertia_diag.z}};
			const M3 R_li = mul(R_lb, rpy(l.com_rpy.x, l.com_rpy.y, l.com_rpy.z));	 // inertial frame in the body frame
			const M3 Ic = mul(mul(R_li, Id), transpose(R_li));
			am += l.mass;
			ah = add(ah, V3{l.mass * c.x, l.mass * c.y, l.mass * c.z});
			add_second_moment(aI, Ic, l.mass, c);
		}
	}
	if (body >= 0) finish(body);
	d.n = body + 1;
	return b;
}

const double H = 1.57079632679;	 // the literal the URDFs use, not pi/2

std::vector<Link> panda_links() {
	return {
		{"link0", FIXED, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, 0, 0, 0, 0, 4.0, {0, 0, 0.05}, {0.4, 0.4, 0.4}},
		{"link1", REVOLUTE, {0, 0, 0.333}, {0, 0, 0}, {0, 0, 1}, -2.8973, 2.8973, 2.1750, 87, 3.0, {0, 0, -0.07}, {0.3, 0.3, 0.3}},
		{"link2", REVOLUTE, {0, 0, 0}, {-H, 0, 0}, {0, 0, 1}, -1.7628, 1.7628, 2.1750, 87, 3.0, {0, -0.1, 0}, {0.3, 0.3, 0.3}},
		{"link3", REVOLUTE, {0, -0.316, 0}, {H, 0, 0}, {0, 0, 1}, -2.8973, 2.8973, 2.1750, 87, 2.0, {0.04, 0, -0.05}, {0.2, 0.2, 0.2}},
		{"link4", REVOLUTE, {0.0825, 0, 0}, {H, 0, 0}, {0, 0, 1}, -3.0718, -0.0698, 2.1750, 87, 2.0, {-0.04, 0.05, 0}, {0.2, 0.2, 0.2}},
		{"link5", REVOLUTE, {-0.0825, 0.384, 0}, {-H, 0, 0}, {0, 0, 1}, -2.8973, 2.8973, 2.6100, 12, 2.0, {0, 0, -0.15}, {0.2, 0.2, 0.2}},
		{"link6", REVOLUTE, {0, 0, 0}, {H, 0, 0}, {0, 0, 1}, -0.0175, 3.7525, 2.6100, 12, 1.5, {0.06, 0, 0}, {0.1, 0.1, 0.1}},
		{"link7", REVOLUTE, {0.088, 0, 0}, {H, 0, 0}, {0, 0, 1}, -2.8973, 2.8973, 2.6100, 12, 1.8, {0, 0, 0.17}, {0.09, 0.05, 0.07}},
		{"end-effector", FIXED, {0, 0, 0.15}, {0, 0, 0}, {0, 0, 1}, 0, 0, 0, 0, 0.2, {0, 0, 0}, {0.01, 0.01, 0.01}},
	};
}

std::vector<Link> sliding_links() {
	std::vector<Link> l = panda_links();
	// joint0 origin literal in the URDF is malformed ("0 0 0.-75"); atof() gives 0
	l[0] = {"link0", PRISMATIC, {0, 0, 0}, {0, 0, 0}, {0, 1, 0}, -1, 1, 2.0, 150, 4.0, {0, 0, 0.05}, {0.4, 0.4, 0.4}};
	l.insert(l.begin(), Link{"slider_link", FIXED, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, 0, 0, 0, 0, 4.0, {0, 0, 0.05}, {0.4, 0.4, 0.4}});
	return l;
}

std::vector<Link> rrrr_links() {
	const V3 I{0.084167, 0.083467, 0.000967};
	std::vector<Link> l = {{"link0", FIXED, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, 0, 0, 0, 0, 1.0, {0, 0, 0}, I},
						   {"link1", REVOLUTE, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, -2.9, 2.9, 1.7104, 176, 1.0, {0.25, 0, 0}, I},
						   {"link2", REVOLUTE, {0.5, 0, 0}, {0, 0, 0}, {0, 0, 1}, -2.9, 2.9, 1.7104, 176, 1.0, {0.25, 0, 0}, I},
						   {"link3", REVOLUTE, {0.5, 0, 0}, {0, 0, 0}, {0, 0, 1}, -2.9, 2.9, 1.7104, 176, 1.0, {0.25, 0, 0}, I},
						   {"link4", REVOLUTE, {0.5, 0, 0}, {0, 0, 0}, {0, 0, 1}, -2.9, 2.9, 1.7104, 176, 1.0, {0.25, 0, 0}, I}};
	return l;
}

std::vector<Link> puma_links() {
	const double h = M_PI / 2;
	return {
		{"base", FIXED, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, 0, 0, 0, 0, 10.0, {0, 0, 0.3}, {1.0, 1.0, 0.5}},
		{"shoulder", REVOLUTE, {0, 0, 0.66}, {0, 0, 0}, {0, 0, 1}, -2.79, 2.79, 2.0, 100, 8.0, {0, 0, -0.1}, {0.30, 0.30, 0.35}},
		{"upper_arm", REVOLUTE, {0, 0.15, 0}, {-h, 0, 0}, {0, 0, 1}, -3.92, 0.78, 2.0, 100, 12.0, {0.20, 0, 0.05}, {0.13, 0.52, 0.54}},
		{"forearm", REVOLUTE, {0.4318, 0, 0}, {0, 0, 0}, {0, 0, 1}, -0.78, 3.92, 2.0, 80, 4.8, {0.02, -0.15, 0}, {0.066, 0.0125, 0.086}},
		{"wrist1", REVOLUTE, {0.0203, -0.4331, 0}, {h, 0, 0}, {0, 0, 1}, -1.92, 2.97, 3.0, 20, 0.82, {0, 0, -0.02}, {0.0018, 0.0018, 0.0013}},
		{"wrist2", REVOLUTE, {0, 0, 0}, {-h, 0, 0}, {0, 0, 1}, -1.74, 1.74, 3.0, 20, 0.34, {0, 0, 0}, {0.0003, 0.0003, 0.0004}},
		{"wrist3", REVOLUTE, {0, 0, 0}, {h, 0, 0}, {0, 0, 1}, -4.64, 4.64, 3.0, 20, 0.09, {0, 0, 0.03}, {0.00015, 0.00015, 0.00004}},
		{"end-effector", FIXED, {0, 0, 0.056}, {0, 0, 0}, {0, 0, 1}, 0, 0, 0, 0, 0.05, {0, 0, 0}, {0.00001, 0.00001, 0.00001}},
	};
}

std::map<std::string, Built>& registered() {
	static std::map<std::string, Built> m;
	return m;
}

const Built* lookup(const char* name) {
	if (name) {
		auto r = registered().find(name);
		if (r != registered().end()) return &r->second;
	}
	static const std::map<std::string, Built> models = [] {
		std::map<std::string, Built> m;
		m["panda"] = build(panda_links());
		m["panda_sliding_base"] = build(sliding_links());
		m["rrrr"] = build(rrrr_links());
		m["puma_like"] = build(puma_links());
		return m;
	}();
	if (!name) return nullptr;
	auto it = models.find(name);
	return it == models.end() ? nullptr : &it->second;
}

// ------------------------------------------------------------------------------------------------------------------
// URDF -> model.  Replaces `std::make_shared<SaiModel::SaiModel>(robot_file)`
// (examples/05-using_robot_controller/05-using_robot_controller.cpp:64) for serial chains: the subset of URDF that
// carries the dynamics (link/inertial, joint/parent/child/origin/axis/limit); visuals and collisions are skipped.
struct XmlTag {
	std::string name;
	std::map<std::string, std::string> attr;
	bool closing = false, self_closing = false;
};

// next tag at or after pos; comments, processing instructions and text are skipped
bool next_tag(const std::string& x, size_t& pos, XmlTag& t) {
	for (;;) {
		const size_t lt = x.find('<', pos);
		if (lt == std::string::npos) return false;
		if (x.compare(lt, 4, "<!--") == 0) {
			const size_t e = x.find("-->", lt + 4);
			if (e == std::string::npos) return false;
			pos = e + 3;
			continue;
		}
		if (x.compare(lt, 2, "<?") == 0 || x.compare(lt, 2, "<!") == 0) {
			const size_t e = x.find('>', lt);
			if (e == std::string::npos) return false;
			pos = e + 1;
			continue;
		}
		size_t gt = lt + 1;
		char quote = 0;
		for (; gt < x.size(); gt++) {  // '>' inside a quoted value does not end the tag
			const char c = x[gt];
			if (quote) {
				if (c == quote) quote = 0;
			} else if (c == '"' || c == '\'') {
				quote = c;
			} else if (c == '>') {
				break;
			}
		}
		if (gt >= x.size()) return false;
		std::string body = x.substr(lt + 1, gt - lt - 1);
		pos = gt + 1;
		t = XmlTag();
		if (!body.empty() && body[0] == '/') {
			t.closing = true;
			body = body.substr(1);
		}
		if (!body.empty() && body.back() == '/') {
			t.self_closing = true;
			body.pop_back();
		}
		size_t i = 0;
		auto skip_ws = [&] { while (i < body.size() && std::isspace((unsigned char)body[i])) i++; };
		skip_ws();
		const size_t n0 = i;
		while (i < body.size() && !std::isspace((unsigned char)body[i])) i++;
		t.name = body.substr(n0, i - n0);
		for (;;) {
			skip_ws();
			if (i >= body.size()) break;
			const size_t k0 = i;
			while (i < body.size() && body[i] != '=' && !std::isspace((unsigned char)body[i])) i++;
			const std::string key = body.substr(k0, i - k0);
			skip_ws();
			if (i >= body.size() || body[i] != '=') continue;
			i++;
			skip_ws();
			if (i >= body.size() || (body[i] != '"' && body[i] != '\'')) break;
			const char q = body[i++];
			const size_t v0 = i;
			while (i < body.size() && body[i] != q) i++;
			t.attr[key] = body.substr(v0, i - v0);
			if (i < body.size()) i++;
		}
		return true;
	}
}

// "a b c" the way the URDF readers do it: strtod one number after the other (a malformed literal such as "0.-75"
// yields 0, like atof)
V3 parse_v3(const std::string& s, V3 def) {
	double v[3] = {def.x, def.y, def.z};
	const char* p = s.c_str();
	for (int k = 0; k < 3; k++) {
		char* e = nullptr;
		const double d = std::strtod(p, &e);
		if (e == p) break;
		v[k] = d;
		p = e;
		while (*p && !std::isspace((unsigned char)*p)) p++;	 // drop the unparsed rest of a malformed token
	}
	return V3{v[0], v[1], v[2]};
}
double parse_d(const std::map<std::string, std::string>& a, const char* key, double def) {
	auto it = a.find(key);
	if (it == a.end()) return def;
	char* e = nullptr;
	const double d = std::strtod(it->second.c_str(), &e);
	return e == it->second.c_str() ? def : d;
}

struct UrdfLink {
	std::string name;
	double mass = 0;
	V3 com{0, 0, 0}, com_rpy{0, 0, 0}, idiag{0, 0, 0}, ioff{0, 0, 0};
};
struct UrdfJoint {
	std::string name, type, parent, child;
	V3 xyz{0, 0, 0}, rpy_{0, 0, 0}, axis{1, 0, 0};
	double lower = 0, upper = 0, velocity = 0, effort = 0;
	bool has_limit = false;
};

// returns an empty string on success, else the reason
std::string urdf_to_built(const std::string& xml, Built& out, std::vector<std::string>& name_store) {
	std::vector<UrdfLink> links;
	std::vector<UrdfJoint> joints;
	size_t pos = 0;
	XmlTag t;
	int in_link = -1, in_joint = -1;
	bool in_inertial = false;
	int skip_depth = 0;	 // inside <visual>/<collision>/<transmission>/...: ignore everything
	std::string skip_name;
	bool saw_robot = false;
	while (next_tag(xml, pos, t)) {
		if (skip_depth > 0) {
			if (t.name == skip_name) {
				if (t.closing) skip_depth--;
				else if (!t.self_closing) skip_depth++;
			}
			continue;
		}
		if (t.name == "robot") {
			saw_robot = true;
			continue;
		}
		if (t.closing) {
			if (t.name == "link") in_link = -1;
			else if (t.name == "joint") in_joint = -1;
			else if (t.name == "inertial") in_inertial = false;
			continue;
		}
		if (in_link < 0 && in_joint < 0) {
			if (t.name == "link") {
				UrdfLink l;
				l.name = t.attr["name"];
				links.push_back(l);
				if (!t.self_closing) in_link = (int)links.size() - 1;
			} else if (t.name == "joint") {
				UrdfJoint j;
				j.name = t.attr["name"];
				j.type = t.attr["type"];
				joints.push_back(j);
				if (!t.self_closing) in_joint = (int)joints.size() - 1;
			} else if (!t.self_closing) {  // material, gazebo, transmission ... at robot level
				skip_name = t.name;
				skip_depth = 1;
			}
			continue;
		}
		if (in_link >= 0) {
			UrdfLink& l = links[in_link];
			if (t.name == "inertial") {
				in_inertial = !t.self_closing;
			} else if (in_inertial && t.name == "origin") {
				l.com = parse_v3(t.attr["xyz"], V3{0, 0, 0});
				l.com_rpy = parse_v3(t.attr["rpy"], V3{0, 0, 0});
			} else if (in_inertial && t.name == "mass") {
				l.mass = parse_d(t.attr, "value", 0.0);
			} else if (in_inertial && t.name == "inertia") {
				l.idiag = V3{parse_d(t.attr, "ixx", 0), parse_d(t.attr, "iyy", 0), parse_d(t.attr, "izz", 0)};
				l.ioff = V3{parse_d(t.attr, "ixy", 0), parse_d(t.attr, "ixz", 0), parse_d(t.attr, "iyz", 0)};
			} else if (!in_inertial && !t.self_closing) {  // visual, collision
				skip_name = t.name;
				skip_depth = 1;
			}
			continue;
		}
		UrdfJoint& j = joints[in_joint];
		if (t.name == "parent") j.parent = t.attr["link"];
		else if (t.name == "child") j.child = t.attr["link"];
		else if (t.name == "origin") {
			j.xyz = parse_v3(t.attr["xyz"], V3{0, 0, 0});
			j.rpy_ = parse_v3(t.attr["rpy"], V3{0, 0, 0});
		} else if (t.name == "axis") j.axis = parse_v3(t.attr["xyz"], V3{1, 0, 0});
		else if (t.name == "limit") {
			j.has_limit = true;
			j.lower = parse_d(t.attr, "lower", 0);
			j.upper = parse_d(t.attr, "upper", 0);
			j.velocity = parse_d(t.attr, "velocity", 0);
			j.effort = parse_d(t.attr, "effort", 0);
		} else if (!t.self_closing) {
			skip_name = t.name;
			skip_depth = 1;
		}
	}
	if (!saw_robot || links.empty()) return "no <robot> with links found";
	// the chain: root = the link that is nobody's child; every link must have at most one child joint
	std::map<std::string, int> link_index, child_joint_of, parent_joint_of;
	for (size_t i = 0; i < links.size(); i++) link_index[links[i].name] = (int)i;
	for (size_t k = 0; k < joints.size(); k++) {
		const UrdfJoint& j = joints[k];
		if (!link_index.count(j.parent) || !link_index.count(j.child)) return "joint [" + j.name + "] refers to an unknown link";
		if (child_joint_of.count(j.parent)) return "link [" + j.parent + "] has several children: only serial chains are supported";
		if (parent_joint_of.count(j.child)) return "link [" + j.child + "] has several parents";
		child_joint_of[j.parent] = (int)k;
		parent_joint_of[j.child] = (int)k;
	}
	std::string root;
	for (const UrdfLink& l : links)
		if (!parent_joint_of.count(l.name)) {
			if (!root.empty()) return "several root links: only one serial chain is supported";
			root = l.name;
		}
	if (root.empty()) return "no root link (kinematic loop)";
	name_store.clear();
	name_store.reserve(links.size());
	std::vector<Link> chain;
	std::string cur = root;
	const UrdfJoint* via = nullptr;
	int dof = 0;
	for (size_t guard = 0; guard <= links.size(); guard++) {
		const UrdfLink& ul = links[link_index[cur]];
		name_store.push_back(ul.name);
		Link l{};
		l.name = nullptr;  // fixed up below (name_store may not reallocate: reserved)
		l.jtype = FIXED;
		l.xyz = V3{0, 0, 0};
		l.rpy_ = V3{0, 0, 0};
		l.axis = V3{0, 0, 1};
		if (via) {
			l.xyz = via->xyz;
			l.rpy_ = via->rpy_;
			if (via->type == "revolute" || via->type == "continuous" || via->type == "prismatic") {
				l.jtype = via->type == "prismatic" ? PRISMATIC : REVOLUTE;
				l.axis = via->axis;
				const double big = 1.7976931348623157e308;
				l.lower = (via->type == "continuous" || !via->has_limit) ? -big : via->lower;
				l.upper = (via->type == "continuous" || !via->has_limit) ? big : via->upper;
				l.velocity = via->has_limit ? via->velocity : big;
				l.effort = via->has_limit ? via->effort : big;
				dof++;
			} else if (via->type != "fixed") {
				return "joint [" + via->name + "] has unsupported type [" + via->type + "]";
			}
		}
		l.mass = ul.mass;
		l.com = ul.com;
		l.inertia_diag = ul.idiag;
		l.inertia_off = ul.ioff;
		l.com_rpy = ul.com_rpy;
		chain.push_back(l);
		auto nx = child_joint_of.find(cur);
		if (nx == child_joint_of.end()) break;
		via = &joints[nx->second];
		cur = via->child;
	}
	if (chain.size() != links.size()) return "links outside the chain";
	if (dof < 1 || dof > OSC_MAX_DOF) return "the chain must have between 1 and OSC_MAX_DOF moving joints";
	for (size_t i = 0; i < chain.size(); i++) chain[i].name = name_store[i].c_str();
	out = build(chain);
	return "";
}

std::string& urdf_error() {
	static std::string e;
	return e;
}

}  // namespace

/* SURVEY.md row f-3 */
extern "C" int osc_urdf_register(const char* model_name, const char* urdf_xml) {
	if (!model_name || !urdf_xml || !*model_name) {
		urdf_error() = "null argument";
		return OSC_ERR_INVALID_ARGUMENT;
	}
	static std::map<std::string, std::vector<std::string>> names;	// keeps the link names of registered models alive
	Built b;
	std::vector<std::string> store;
	const std::string why = urdf_to_built(urdf_xml, b, store);
	if (!why.empty()) {
		urdf_error() = why;
		return OSC_ERR_INVALID_ARGUMENT;
	}
	names[model_name] = std::move(store);
	registered()[model_name] = b;
	urdf_error().clear();
	return OSC_OK;
}

extern "C" int osc_urdf_register_file(const char* model_name, const char* path) {
	if (!path) {
		urdf_error() = "null path";
		return OSC_ERR_INVALID_ARGUMENT;
	}
	FILE* f = std::fopen(path, "rb");
	if (!f) {
		urdf_error() = std::string("cannot open [") + path + "]";
		return OSC_ERR_INVALID_ARGUMENT;
	}
	std::string xml;
	char buf[65536];
	size_t n;
	while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) xml.append(buf, n);
	std::fclose(f);
	return osc_urdf_register(model_name, xml.c_str());
}

extern "C" const char* osc_urdf_last_error(void) { return urdf_error().c_str(); }

// ---- ${PREFIX} substitution and world files -------------------------------------------------------
// The reference's examples name their files "${EXAMPLE_15_FOLDER}/world.urdf", "${SAI_MODEL_URDF_FOLDER}/panda/panda_arm.urdf"
// and resolve them with SaiModel::URDF_FOLDERS / SaiModel::ReplaceUrdfPathPrefix
// (examples/01-joint_control/01-joint_control.cpp:40-71); the robot's base pose comes from the <robot><origin> element of the
// world file (sim->getRobotBaseTransform + robot->setTRobotBase, examples/15-haptic_control_impedance_type/...cpp:124-126).
namespace {
std::map<std::string, std::string>& urdf_folders() {
	static std::map<std::string, std::string> m;
	return m;
}
std::string replace_prefixes(const std::string& in, std::string& why) {
	std::string out;
	size_t pos = 0;
	for (;;) {
		const size_t a = in.find("${", pos);
		if (a == std::string::npos) {
			out += in.substr(pos);
			return out;
		}
		const size_t b = in.find('}', a);
		if (b == std::string::npos) {
			why = "unterminated ${ in [" + in + "]";
			return in;
		}
		const std::string key = in.substr(a + 2, b - a - 2);
		auto it = urdf_folders().find(key);
		if (it == urdf_folders().end()) {
			why = "prefix ${" + key + "} is not set (osc_urdf_set_folder)";
			return in;
		}
		out += in.substr(pos, a - pos) + it->second;
		pos = b + 1;
	}
}
bool read_file(const std::string& path, std::string& xml) {
	FILE* f = std::fopen(path.c_str(), "rb");
	if (!f) return false;
	char buf[65536];
	size_t n;
	while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) xml.append(buf, n);
	std::fclose(f);
	return true;
}
}  // namespace

extern "C" int osc_urdf_set_folder(const char* prefix_name, const char* folder) {
	if (!prefix_name || !folder || !*prefix_name) return OSC_ERR_INVALID_ARGUMENT;
	urdf_folders()[prefix_name] = folder;
	return OSC_OK;
}

extern "C" int osc_urdf_replace_path_prefix(const char* path, char* out, int out_capacity) {
	if (!path || !out || out_capacity <= 0) return OSC_ERR_INVALID_ARGUMENT;
	std::string why;
	const std::string r = replace_prefixes(path, why);
	if (!why.empty()) {
		urdf_error() = why;
		return OSC_ERR_INVALID_ARGUMENT;
	}
	if ((int)r.size() + 1 > out_capacity) {
		urdf_error() = "output buffer too small";
		return OSC_ERR_INVALID_ARGUMENT;
	}
	std::memcpy(out, r.c_str(), r.size() + 1);
	return OSC_OK;
}

extern "C" int osc_world_register_robot(const char* world_file, const char* robot_name_in_world, const char* model_name, double gravity_out[3]) {
	if (!world_file || !robot_name_in_world || !model_name) {
		urdf_error() = "null argument";
		return OSC_ERR_INVALID_ARGUMENT;
	}
	std::string why;
	const std::string wpath = replace_prefixes(world_file, why);
	std::string xml;
	if (why.empty() && !read_file(wpath, xml)) why = "cannot open [" + wpath + "]";
	if (!why.empty()) {
		urdf_error() = why;
		return OSC_ERR_INVALID_ARGUMENT;
	}
	size_t pos = 0;
	XmlTag t;
	bool in_robot = false, found = false;
	int depth_in_robot = 0;
	std::string dir, file;
	V3 xyz{0, 0, 0}, rpy_{0, 0, 0}, gravity{0, 0, -9.81};
	while (next_tag(xml, pos, t)) {
		if (!in_robot) {
			if (t.name == "world" && !t.closing && t.attr.count("gravity")) gravity = parse_v3(t.attr["gravity"], gravity);
			if (t.name == "robot" && !t.closing && t.attr.count("name") && t.attr["name"] == robot_name_in_world) {
				in_robot = true;
				found = true;
				depth_in_robot = 0;
			}
			continue;
		}
		if (t.closing) {
			if (depth_in_robot == 0 && t.name == "robot") break;
			depth_in_robot--;
			continue;
		}
		if (depth_in_robot == 0 && t.name == "model") {
			dir = t.attr.count("dir") ? t.attr["dir"] : "";
			file = t.attr.count("path") ? t.attr["path"] : "";
		} else if (depth_in_robot == 0 && t.name == "origin") {
			if (t.attr.count("xyz")) xyz = parse_v3(t.attr["xyz"], xyz);
			if (t.attr.count("rpy")) rpy_ = parse_v3(t.attr["rpy"], rpy_);
		}
		if (!t.self_closing) depth_in_robot++;
	}
	if (!found || file.empty()) {
		urdf_error() = std::string("robot [") + robot_name_in_world + "] with a <model path=...> not found in [" + wpath + "]";
		return OSC_ERR_INVALID_ARGUMENT;
	}
	const std::string upath = replace_prefixes(dir.empty() ? file : dir + "/" + file, why);
	if (!why.empty()) {
		urdf_error() = why;
		return OSC_ERR_INVALID_ARGUMENT;
	}
	const int rc = osc_urdf_register_file(model_name, upath.c_str());
	if (rc != OSC_OK) return rc;
	// SaiModel::setTRobotBase(T_world_robot): R = Rz(yaw) Ry(pitch) Rx(roll) of the <origin> (URDF convention)
	Built& b = registered()[model_name];
	const M3 R = rpy(rpy_.x, rpy_.y, rpy_.z);
	for (int i = 0; i < 9; i++) b.desc.R_world_base[i] = R.m[i];
	b.desc.t_world_base[0] = xyz.x;
	b.desc.t_world_base[1] = xyz.y;
	b.desc.t_world_base[2] = xyz.z;
	if (gravity_out) {
		gravity_out[0] = gravity.x;
		gravity_out[1] = gravity.y;
		gravity_out[2] = gravity.z;
	}
	return OSC_OK;
}

extern "C" int osc_builtin_model(const char* robot_name, osc_model_desc* out) {
	const Built* b = lookup(robot_name);
	if (!b || !out) return OSC_ERR_INVALID_ARGUMENT;
	*out = b->desc;
	return OSC_OK;
}

extern "C" int osc_builtin_link(const char* robot_name, const char* link_name, osc_link_frame* out) {
	const Built* b = lookup(robot_name);
	if (!b || !out || !link_name) return OSC_ERR_INVALID_ARGUMENT;
	auto it = b->frames.find(link_name);
	if (it == b->frames.end()) return OSC_ERR_INVALID_ARGUMENT;
	*out = it->second;
	return OSC_OK;
}
