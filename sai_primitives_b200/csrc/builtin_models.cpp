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
#include <cmath>
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
			M3 Id{{l.inertia_diag.x, 0, 0, 0, l.inertia_diag.y, 0, 0, 0, l.inertia_diag.z}};
			const M3 Ic = mul(mul(R_lb, Id), transpose(R_lb));
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

const Built* lookup(const char* name) {
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

}  // namespace

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
