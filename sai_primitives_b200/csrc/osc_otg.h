// Online trajectory generation for the tasks' internal OTG, acceleration-limited and phase-synchronised:
// what the reference obtains from its vendored Ruckig 0.9 with max_jerk = infinity (OTG_joints::disableJerkLimits,
// src/helper_modules/OTG_joints.cpp:89-92 -- the default of JointTask.h:38-42 and MotionForceTask.h:67-74).
//
// Restated from Ruckig's second-order position interface (files under /root/reference/ruckig):
//   src/ruckig/position-second-step1.cpp      extremal profiles of one degree of freedom        -> step1()
//   include/ruckig/block.hpp:62-133           blocked duration intervals                        -> calculate_block()
//   include/ruckig/calculator_target.hpp:123-202   synchronised duration                         -> synchronize()
//   include/ruckig/calculator_target.hpp:43-119, :366-433   phase synchronisation (collinear inputs)
//   src/ruckig/position-second-step2.cpp      profile of given duration                         -> step2()
//   src/ruckig/brake.cpp:82-102, include/ruckig/brake.hpp:64-74   braking pre-trajectory
//   include/ruckig/profile.hpp:306-361        profile integration and limit checks             -> Profile::check()
//   include/ruckig/trajectory.hpp:64-143      state at a given time                              -> Trajectory::at_time()
// Plain C++ that compiles for host and device: the device kernels (osc_otg_kernels.cuh) call it per robot, and a host build of
// the same header is checked against the reference's own Ruckig on the CPU (tests/test_otg_core.py).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OTG_HD __host__ __device__
#else
#define OTG_HD
#endif

namespace otg {

constexpr double kEps = 2.220446049250313e-16;	// DBL_EPSILON
constexpr double kInf = __builtin_huge_val();
enum Result : int { WORKING = 0, FINISHED = 1, ERROR_INVALID_INPUT = -100, ERROR_EXECUTION_TIME = -110, ERROR_SYNCHRONIZATION = -111, ERROR_DURATION = -101 };

// include/ruckig/profile.hpp, the members the second-order interface uses
struct Profile {
	double t[7], t_sum[7], a[8], v[8], p[8];
	double pf, vf, af;
	double brake_duration, brake_t0, brake_a0, brake_p0, brake_v0;
	int direction;	// 0 UP, 1 DOWN

	OTG_HD void set_boundary(const Profile& o) {  // profile.hpp:285-294
		a[0] = o.a[0];
		v[0] = o.v[0];
		p[0] = o.p[0];
		af = o.af;
		vf = o.vf;
		pf = o.pf;
		brake_duration = o.brake_duration;
		brake_t0 = o.brake_t0;
		brake_a0 = o.brake_a0;
		brake_p0 = o.brake_p0;
		brake_v0 = o.brake_v0;
	}
	// profile.hpp:306-350 (control signs UDDU: the only ones the second-order interface produces)
	OTG_HD bool check(double aUp, double aDown, double vMax, double vMin) {
		if (t[0] < 0) return false;
		t_sum[0] = t[0];
		for (int i = 0; i < 6; ++i) {
			if (t[i + 1] < 0) return false;
			t_sum[i + 1] = t_sum[i] + t[i + 1];
		}
		if (t_sum[6] > 1e12) return false;
		a[0] = (t[0] > 0 ? aUp : 0);
		a[1] = 0;
		a[2] = (t[2] > 0 ? aDown : 0);
		a[3] = 0;
		a[4] = (t[4] > 0 ? aDown : 0);
		a[5] = 0;
		a[6] = (t[6] > 0 ? aUp : 0);
		a[7] = af;
		direction = (vMax > 0) ? 0 : 1;
		const double vUppLim = (direction == 0 ? vMax : vMin) + 1e-12;
		const double vLowLim = (direction == 0 ? vMin : vMax) - 1e-12;
		for (int i = 0; i < 7; ++i) {
			v[i + 1] = v[i] + t[i] * a[i];
			p[i + 1] = p[i] + t[i] * (v[i] + t[i] * a[i] / 2);
		}
		return fabs(p[7] - pf) < 1e-8 && fabs(v[7] - vf) < 1e-8 && v[2] <= vUppLim && v[3] <= vUppLim && v[4] <= vUppLim && v[5] <= vUppLim &&
			   v[6] <= vUppLim && v[2] >= vLowLim && v[3] >= vLowLim && v[4] >= vLowLim && v[5] >= vLowLim && v[6] >= vLowLim;
	}
	// profile.hpp:358-361
	OTG_HD bool check_with_timing(double aUp, double aDown, double vMax, double vMin, double aMax, double aMin) {
		return (aMin - 1e-12 < aUp) && (aUp < aMax + 1e-12) && (aMin - 1e-12 < aDown) && (aDown < aMax + 1e-12) && check(aUp, aDown, vMax, vMin);
	}
	OTG_HD double duration() const { return t_sum[6] + brake_duration; }
};

// include/ruckig/block.hpp
struct Block {
	Profile p_min;
	double t_min;
	bool has_a;	 // at most one blocked interval with three valid profiles (position.hpp:105-107)
	double a_left, a_right;
	Profile a_profile;
	OTG_HD void set_min_profile(const Profile& pr) {
		p_min = pr;
		t_min = pr.duration();
		has_a = false;
	}
	OTG_HD void set_interval(const Profile& l, const Profile& r) {	// Interval(profile_left, profile_right), block.hpp:31-43
		const double ld = l.duration(), rd = r.duration();
		has_a = true;
		if (ld < rd) {
			a_left = ld;
			a_right = rd;
			a_profile = r;
		} else {
			a_left = rd;
			a_right = ld;
			a_profile = l;
		}
	}
	OTG_HD bool is_blocked(double t) const { return (t < t_min) || (has_a && a_left < t && t < a_right); }
};

// block.hpp:62-133 for at most three valid profiles
OTG_HD inline bool calculate_block(Block& block, const Profile* valid, int count) {
	if (count == 1) {
		block.set_min_profile(valid[0]);
		return true;
	}
	if (count == 2) {
		if (fabs(valid[0].t_sum[6] - valid[1].t_sum[6]) < 8 * kEps) {
			block.set_min_profile(valid[0]);
			return true;
		}
		const int idx_min = (valid[0].t_sum[6] < valid[1].t_sum[6]) ? 0 : 1;
		const int idx_else = (idx_min + 1) % 2;
		block.set_min_profile(valid[idx_min]);
		block.set_interval(valid[idx_min], valid[idx_else]);
		return true;
	}
	if (count != 3) return false;
	int idx_min = 0;
	for (int i = 1; i < 3; i++)
		if (valid[i].t_sum[6] < valid[idx_min].t_sum[6]) idx_min = i;
	block.set_min_profile(valid[idx_min]);
	block.set_interval(valid[(idx_min + 1) % 3], valid[(idx_min + 2) % 3]);
	return true;
}

// src/ruckig/position-second-step1.cpp
struct Step1 {
	double v0, vf, vMax_, vMin_, aMax_, aMin_, pd;
	Profile valid[3];
	int count;

	OTG_HD void add_profile() {	 // position.hpp:113-117
		if (count < 2) valid[count + 1].set_boundary(valid[count]);
		count++;
	}
	OTG_HD void time_acc0(double vMax, double vMin, double aMax, double aMin) {	 // :10-24
		if (count >= 3) return;
		Profile& pr = valid[count];
		pr.t[0] = (-v0 + vMax) / aMax;
		pr.t[1] = (aMin * v0 * v0 - aMax * vf * vf) / (2 * aMax * aMin * vMax) + vMax * (aMax - aMin) / (2 * aMax * aMin) + pd / vMax;
		pr.t[2] = (vf - vMax) / aMin;
		pr.t[3] = pr.t[4] = pr.t[5] = pr.t[6] = 0;
		if (pr.check(aMax, aMin, vMax, vMin)) add_profile();
	}
	OTG_HD void time_none(double vMax, double vMin, double aMax, double aMin, bool return_after_found) {  // :26-63
		double h1 = (aMax * vf * vf - aMin * v0 * v0 - 2 * aMax * aMin * pd) / (aMax - aMin);
		if (h1 >= 0.0) {
			h1 = sqrt(h1);
			if (count < 3) {
				Profile& pr = valid[count];
				pr.t[3] = pr.t[4] = pr.t[5] = pr.t[6] = 0;
				pr.t[0] = -(v0 + h1) / aMax;
				pr.t[1] = 0;
				pr.t[2] = (vf + h1) / aMin;
				if (pr.check(aMax, aMin, vMax, vMin)) {
					add_profile();
					if (return_after_found) return;
				}
			}
			if (count < 3) {
				Profile& pr = valid[count];
				pr.t[3] = pr.t[4] = pr.t[5] = pr.t[6] = 0;
				pr.t[0] = (-v0 + h1) / aMax;
				pr.t[1] = 0;
				pr.t[2] = (vf - h1) / aMin;
				if (pr.check(aMax, aMin, vMax, vMin)) add_profile();
			}
		}
	}
	// :101-137 (limits are strictly positive here: the wrappers reject zero limits, OTG_joints.cpp:50-54, :66-70)
	OTG_HD bool get_profile(const Profile& input, Block& block) {
		count = 0;
		valid[0].set_boundary(input);
		if (fabs(vf) < kEps) {
			const double vMax = (pd >= 0) ? vMax_ : vMin_, vMin = (pd >= 0) ? vMin_ : vMax_;
			const double aMax = (pd >= 0) ? aMax_ : aMin_, aMin = (pd >= 0) ? aMin_ : aMax_;
			time_none(vMax, vMin, aMax, aMin, true);
			if (count == 0) time_acc0(vMax, vMin, aMax, aMin);
			if (count == 0) time_none(vMin, vMax, aMin, aMax, true);
			if (count == 0) time_acc0(vMin, vMax, aMin, aMax);
		} else {
			time_none(vMax_, vMin_, aMax_, aMin_, false);
			time_none(vMin_, vMax_, aMin_, aMax_, false);
			time_acc0(vMax_, vMin_, aMax_, aMin_);
			time_acc0(vMin_, vMax_, aMin_, aMax_);
		}
		return calculate_block(block, valid, count);
	}
};

// src/ruckig/position-second-step2.cpp
struct Step2 {
	double v0, tf, vf, vMax_, vMin_, aMax_, aMin_, pd, vd;
	OTG_HD bool time_acc0(Profile& pr, double vMax, double vMin, double aMax, double aMin) {  // :14-68
		{
			const double h1 = sqrt((2 * aMax * (pd - tf * vf) - 2 * aMin * (pd - tf * v0) + vd * vd) / (aMax * aMin) + tf * tf);
			pr.t[0] = (aMax * vd - aMax * aMin * (tf - h1)) / (aMax * (aMax - aMin));
			pr.t[1] = h1;
			pr.t[2] = tf - (pr.t[0] + h1);
			pr.t[3] = pr.t[4] = pr.t[5] = pr.t[6] = 0;
			if (pr.check(aMax, aMin, vMax, vMin)) {
				pr.pf = pr.p[7];
				return true;
			}
		}
		{
			const double h1 = (-vd + aMax * tf);
			pr.t[0] = -vd * vd / (2 * aMax * h1) + (pd - v0 * tf) / h1;
			pr.t[1] = -vd / aMax + tf;
			pr.t[2] = pr.t[3] = pr.t[4] = pr.t[5] = 0;
			pr.t[6] = tf - (pr.t[0] + pr.t[1]);
			if (pr.check(aMax, aMin, vMax, vMin)) {
				pr.pf = pr.p[7];
				return true;
			}
		}
		{
			pr.t[0] = 0;
			pr.t[1] = -vd / aMax + tf;
			pr.t[2] = pr.t[3] = pr.t[4] = pr.t[5] = 0;
			pr.t[6] = vd / aMax;
			if (pr.check(aMax, aMin, vMax, vMin)) {
				pr.pf = pr.p[7];
				return true;
			}
		}
		return false;
	}
	OTG_HD bool time_none(Profile& pr, double vMax, double vMin, double aMax, double aMin) {  // :70-107
		if (fabs(v0) < kEps && fabs(vf) < kEps && fabs(pd) < kEps) {
			pr.t[0] = 0;
			pr.t[1] = tf;
			pr.t[2] = pr.t[3] = pr.t[4] = pr.t[5] = pr.t[6] = 0;
			if (pr.check(aMax, aMin, vMax, vMin)) {
				pr.pf = pr.p[7];
				return true;
			}
		}
		{
			const double h1 = 2 * (vf * tf - pd);
			pr.t[0] = h1 / vd;
			pr.t[1] = tf - pr.t[0];
			pr.t[2] = pr.t[3] = pr.t[4] = pr.t[5] = pr.t[6] = 0;
			const double af = vd * vd / h1;
			if ((aMin - 1e-12 < af) && (af < aMax + 1e-12) && pr.check(af, -af, vMax, vMin)) {
				pr.pf = pr.p[7];
				return true;
			}
		}
		return false;
	}
	OTG_HD bool check_all(Profile& pr, double vMax, double vMin, double aMax, double aMin) {
		return time_acc0(pr, vMax, vMin, aMax, aMin) || time_none(pr, vMax, vMin, aMax, aMin);
	}
	OTG_HD bool get_profile(Profile& pr) {	// :109-117
		if (pd > 0) return check_all(pr, vMax_, vMin_, aMax_, aMin_) || check_all(pr, vMin_, vMax_, aMin_, aMax_);
		return check_all(pr, vMin_, vMax_, aMin_, aMax_) || check_all(pr, vMax_, vMin_, aMax_, aMin_);
	}
};

// What a trajectory needs at sampling time, per degree of freedom (20 doubles)
struct Section {
	double t[7], a[7];
	double p0, v0;	// state after the braking pre-trajectory
	double brake_duration, brake_a0, brake_p0, brake_v0;
	OTG_HD void from(const Profile& pr) {
		for (int i = 0; i < 7; i++) {
			t[i] = pr.t[i];
			a[i] = pr.a[i];
		}
		p0 = pr.p[0];
		v0 = pr.v[0];
		brake_duration = pr.brake_duration;
		brake_a0 = pr.brake_a0;
		brake_p0 = pr.brake_p0;
		brake_v0 = pr.brake_v0;
	}
	// trajectory.hpp:64-143 with one section and jerk 0; af = target acceleration = 0 in the wrappers
	OTG_HD void at_time(double time, double duration, double& pos, double& vel, double& acc) const {
		double ts[7], vv[8], pp[8];
		ts[0] = t[0];
		for (int i = 0; i < 6; i++) ts[i + 1] = ts[i] + t[i + 1];
		vv[0] = v0;
		pp[0] = p0;
		for (int i = 0; i < 7; i++) {
			vv[i + 1] = vv[i] + t[i] * a[i];
			pp[i + 1] = pp[i] + t[i] * (vv[i] + t[i] * a[i] / 2);
		}
		double td, ps, vs, as;
		if (time >= duration) {
			td = time - (brake_duration + ts[6]);
			ps = pp[7];
			vs = vv[7];
			as = 0.0;
		} else {
			td = time;
			bool done = false;
			if (brake_duration > 0) {
				if (td < brake_duration) {
					ps = brake_p0;
					vs = brake_v0;
					as = brake_a0;
					done = true;
				} else {
					td -= brake_duration;
				}
			}
			if (!done) {
				if (td >= ts[6]) {
					td -= ts[6];
					ps = pp[7];
					vs = vv[7];
					as = 0.0;
				} else {
					int idx = 0;
					while (idx < 7 && !(td < ts[idx])) idx++;	 // std::upper_bound
					if (idx > 0) td -= ts[idx - 1];
					ps = pp[idx];
					vs = vv[idx];
					as = a[idx];
				}
			}
		}
		pos = ps + td * (vs + td * (as / 2));	// utils.hpp:44-50 with j = 0
		vel = vs + td * as;
		acc = as;
	}
};

// calculator_target.hpp:205-520 for the configuration the reference's wrappers use: position interface, every degree of
// freedom enabled, no minimum duration, continuous durations, max_jerk = inf, Synchronization::Phase, min limits = -max.
// D = compile-time bound on the degrees of freedom.  Returns WORKING or an error code; on success sec[dof] and duration.
template <int D>
struct Calculator {
	Block blocks[D];
	Profile prof[D];

	OTG_HD int calculate(int dofs, const double* p0, const double* v0, const double* a0, const double* pf, const double* vf, const double* vmax,
						 const double* amax, Section* sec, double& duration) {
		// InputParameter::validate (input_parameter.hpp:155-330) with check_target_state_within_limits
		for (int d = 0; d < dofs; d++) {
			if (isnan(p0[d]) || isnan(v0[d]) || isnan(a0[d]) || isnan(pf[d]) || isnan(vf[d])) return ERROR_INVALID_INPUT;
			if (vf[d] > vmax[d] || vf[d] < -vmax[d]) return ERROR_INVALID_INPUT;
		}
		for (int d = 0; d < dofs; d++) {
			Profile& p = prof[d];
			const double vMax = vmax[d], vMin = -vmax[d], aMax = amax[d], aMin = -amax[d];
			// brake.cpp:82-102
			p.brake_t0 = 0.0;
			p.brake_a0 = 0.0;
			if (v0[d] > vMax) {
				p.brake_a0 = aMin;
				p.brake_t0 = (vMax - v0[d]) / aMin + kEps;
			} else if (v0[d] < vMin) {
				p.brake_a0 = aMax;
				p.brake_t0 = (vMin - v0[d]) / aMax + kEps;
			}
			p.a[0] = a0[d];
			p.v[0] = v0[d];
			p.p[0] = p0[d];
			p.pf = pf[d];
			p.vf = vf[d];
			p.af = 0.0;
			// brake.hpp:64-74
			if (p.brake_t0 <= 0.0) {
				p.brake_duration = 0.0;
				p.brake_p0 = p.brake_v0 = 0.0;
			} else {
				p.brake_duration = p.brake_t0;
				p.brake_p0 = p.p[0];
				p.brake_v0 = p.v[0];
				const double tb = p.brake_t0, ps = p.p[0], vs = p.v[0];
				p.p[0] = ps + tb * (vs + tb * (p.brake_a0 / 2));
				p.v[0] = vs + tb * p.brake_a0;
				p.a[0] = p.brake_a0;
			}
			Step1 s1;
			s1.v0 = p.v[0];
			s1.vf = p.vf;
			s1.vMax_ = vMax;
			s1.vMin_ = vMin;
			s1.aMax_ = aMax;
			s1.aMin_ = aMin;
			s1.pd = p.pf - p.p[0];
			if (!s1.get_profile(p, blocks[d])) return ERROR_EXECUTION_TIME;
		}
		int limiting = -1;
		if (dofs == 1) {
			duration = blocks[0].t_min;
			prof[0] = blocks[0].p_min;
			sec[0].from(prof[0]);
			return WORKING;
		}
		// synchronize (:123-202): candidates are the t_min and the right ends of the blocked intervals, tested in ascending order
		// starting from the largest t_min
		{
			double cand[2 * D];
			int order[2 * D];
			bool any_interval = false;
			for (int d = 0; d < dofs; d++) {
				cand[d] = blocks[d].t_min;
				cand[dofs + d] = blocks[d].has_a ? blocks[d].a_right : kInf;
				any_interval = any_interval || blocks[d].has_a;
			}
			const int n_cand = any_interval ? 2 * dofs : dofs;	// (the third group and the optional t_min are infinite here)
			for (int i = 0; i < n_cand; i++) order[i] = i;
			for (int i = 1; i < n_cand; i++) {	// insertion sort by value (std::sort; ties do not matter for the result)
				const int key = order[i];
				int j = i - 1;
				while (j >= 0 && cand[order[j]] > cand[key]) {
					order[j + 1] = order[j];
					j--;
				}
				order[j + 1] = key;
			}
			bool found = false;
			for (int k = dofs - 1; k < n_cand && !found; k++) {
				const double tc = cand[order[k]];
				bool blocked = false;
				for (int d = 0; d < dofs; d++)
					if (blocks[d].is_blocked(tc)) {
						blocked = true;
						break;
					}
				if (blocked || tc < 0.0 || isinf(tc)) continue;
				duration = tc;
				limiting = order[k] % dofs;
				prof[limiting] = (order[k] / dofs == 0) ? blocks[limiting].p_min : blocks[limiting].a_profile;
				found = true;
			}
			if (!found) return ERROR_SYNCHRONIZATION;
		}
		if (duration > 7.6e3) return ERROR_DURATION;
		if (duration == 0.0) {
			for (int d = 0; d < dofs; d++) {
				prof[d] = blocks[d].p_min;
				sec[d].from(prof[d]);
			}
			return WORKING;
		}
		// phase synchronisation (:43-119, :366-433)
		{
			double pd[D], ctrl[D];
			for (int d = 0; d < dofs; d++) pd[d] = pf[d] - p0[d];
			const double* scale_vector = nullptr;
			int scale_dof = -1;
			for (int d = 0; d < dofs; d++) {
				if (fabs(pd[d]) > kEps) { scale_vector = pd; scale_dof = d; break; }
				if (fabs(v0[d]) > kEps) { scale_vector = v0; scale_dof = d; break; }
				if (fabs(a0[d]) > kEps) { scale_vector = a0; scale_dof = d; break; }
				if (fabs(vf[d]) > kEps) { scale_vector = vf; scale_dof = d; break; }
			}
			bool collinear = scale_dof >= 0;
			if (collinear) {
				const double scale = scale_vector[scale_dof];
				const double pd_scale = pd[scale_dof] / scale, v0_scale = v0[scale_dof] / scale, vf_scale = vf[scale_dof] / scale,
							 a0_scale = a0[scale_dof] / scale;
				const double scale_limiting = scale_vector[limiting];
				const double control_limiting = (prof[limiting].direction == 0) ? amax[limiting] : -amax[limiting];
				for (int d = 0; d < dofs; d++) {
					const double cs = scale_vector[d];
					if (fabs(pd[d] - pd_scale * cs) > kEps || fabs(v0[d] - v0_scale * cs) > kEps || fabs(a0[d] - a0_scale * cs) > kEps ||
						fabs(vf[d] - vf_scale * cs) > kEps) {
						collinear = false;
						break;
					}
					ctrl[d] = control_limiting * cs / scale_limiting;
				}
			}
			if (collinear) {
				bool ok = true;
				for (int d = 0; d < dofs; d++) {
					if (d == limiting) continue;
					Profile& p = prof[d];
					for (int i = 0; i < 7; i++) p.t[i] = prof[limiting].t[i];
					ok = p.check_with_timing(ctrl[d], -ctrl[d], vmax[d], -vmax[d], amax[d], -amax[d]) && ok;
				}
				if (ok) {
					for (int d = 0; d < dofs; d++) sec[d].from(prof[d]);
					return WORKING;
				}
			}
		}
		// time synchronisation (:436-507)
		for (int d = 0; d < dofs; d++) {
			if (d == limiting) continue;
			Profile& p = prof[d];
			const double t_profile = duration - p.brake_duration;
			if (fabs(t_profile - blocks[d].t_min) < 2 * kEps) {
				p = blocks[d].p_min;
				continue;
			}
			if (blocks[d].has_a && fabs(t_profile - blocks[d].a_right) < 2 * kEps) {
				p = blocks[d].a_profile;
				continue;
			}
			Step2 s2;
			s2.tf = t_profile;
			s2.v0 = p.v[0];
			s2.vf = p.vf;
			s2.vMax_ = vmax[d];
			s2.vMin_ = -vmax[d];
			s2.aMax_ = amax[d];
			s2.aMin_ = -amax[d];
			s2.pd = p.pf - p.p[0];
			s2.vd = p.vf - p.v[0];
			if (!s2.get_profile(p)) return ERROR_SYNCHRONIZATION;
		}
		for (int d = 0; d < dofs; d++) sec[d].from(prof[d]);
		return WORKING;
	}
};

// Eigen's isApprox for vectors (default precision 1e-12): ||a - b||^2 <= prec^2 min(||a||^2, ||b||^2)
OTG_HD inline bool is_approx(const double* a, const double* b, int n, double prec) {
	double d2 = 0, na = 0, nb = 0;
	for (int i = 0; i < n; i++) {
		d2 += (a[i] - b[i]) * (a[i] - b[i]);
		na += a[i] * a[i];
		nb += b[i] * b[i];
	}
	return d2 <= prec * prec * (na < nb ? na : nb);
}

enum OtgFlags : int {
	OTG_GOAL_REACHED = 1,	// OTG_joints::_goal_reached
	OTG_DIRTY = 2,			// the wrapper's _input differs from Ruckig's current_input: the next update recalculates (ruckig.hpp:198-208)
	OTG_ERROR = 4,			// last update hit the error branch (OTG_joints.cpp:141-149)
	OTG_BAD_FINISH = 8,		// finished with a non-zero velocity: the reference then calls setGoalPosition with an empty vector and throws
};

// src/helper_modules/OTG_joints.cpp for K degrees of freedom, acceleration-limited (jerk limits disabled)
template <int K>
struct JointsOtg {
	double target_pos[K], target_vel[K];			 // _input.target_position / target_velocity
	double in_pos[K], in_vel[K], in_acc[K];		 // _input.current_*
	double out_pos[K], out_vel[K], out_acc[K];	 // _output.new_*
	double time, duration;						 // _output.time, trajectory duration
	Section sec[K];
	int flags;

	// OTG_joints::setGoalPositionAndVelocity (:99-116)
	OTG_HD void set_goal(int k, const double* goal_pos, const double* goal_vel) {
		if (is_approx(goal_pos, target_pos, k, 1e-12) && is_approx(goal_vel, target_vel, k, 1e-12)) return;
		flags &= ~OTG_GOAL_REACHED;
		flags |= OTG_DIRTY;
		for (int i = 0; i < k; i++) {
			target_pos[i] = goal_pos[i];
			target_vel[i] = goal_vel[i];
		}
	}
	// OTG_joints::reInitialize (:28-41)
	OTG_HD void reinitialize(int k, const double* pos) {
		double zero[K];
		for (int i = 0; i < k; i++) zero[i] = 0.0;
		set_goal(k, pos, zero);
		for (int i = 0; i < k; i++) {
			out_pos[i] = in_pos[i] = pos[i];
			out_vel[i] = in_vel[i] = 0.0;
			out_acc[i] = in_acc[i] = 0.0;
		}
		flags |= OTG_DIRTY;
	}
	// OTG_joints::update (:118-150) around Ruckig::update (ruckig.hpp:186-221); returns true when the trajectory was recalculated
	OTG_HD bool update(int k, double dt, const double* vmax, const double* amax, Calculator<K>* calc) {
		if (flags & OTG_GOAL_REACHED) return false;
		bool recalculated = false;
		flags &= ~(OTG_ERROR | OTG_BAD_FINISH);
		int result = WORKING;
		if (flags & OTG_DIRTY) {
			double dur = 0.0;
			result = calc->calculate(k, in_pos, in_vel, in_acc, target_pos, target_vel, vmax, amax, sec, dur);
			if (result == WORKING) {
				duration = dur;
				time = 0.0;
				flags &= ~OTG_DIRTY;
				recalculated = true;
			}
		}
		if (result != WORKING) {
			// error: the previous output is kept, the trajectory restarts from rest at the next update
			flags |= OTG_ERROR | OTG_DIRTY;
			for (int i = 0; i < k; i++) in_vel[i] = in_acc[i] = 0.0;
			return false;
		}
		time += dt;
		for (int i = 0; i < k; i++) sec[i].at_time(time, duration, out_pos[i], out_vel[i], out_acc[i]);
		if (time > duration) {	// Result::Finished: _input keeps the state of the previous step (no pass_to_input, :128-135)
			double n2 = 0.0;
			for (int i = 0; i < k; i++) n2 += out_vel[i] * out_vel[i];
			if (sqrt(n2) < 1e-3)
				flags |= OTG_GOAL_REACHED;
			else
				flags |= OTG_BAD_FINISH;
			flags |= OTG_DIRTY;
			return recalculated;
		}
		for (int i = 0; i < k; i++) {  // Result::Working: _output.pass_to_input(_input)
			in_pos[i] = out_pos[i];
			in_vel[i] = out_vel[i];
			in_acc[i] = out_acc[i];
		}
		return recalculated;
	}
};

// ---- small 3 x 3 helpers (row-major)
OTG_HD inline void m3_mul(const double* A, const double* B, double* C) {
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
OTG_HD inline void m3t_mul(const double* A, const double* B, double* C) {  // A^T B
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
OTG_HD inline void m3_vec(const double* A, const double* v, double* o) {
	for (int i = 0; i < 3; i++) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
OTG_HD inline void m3t_vec(const double* A, const double* v, double* o) {
	for (int i = 0; i < 3; i++) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
// Eigen::AngleAxisd(Matrix3d): rotation matrix -> quaternion (largest-component branch) -> angle in [0, pi] times unit axis
OTG_HD inline void rotation_vector(const double* m, double* rv) {
	double w, x, y, z;
	double t = m[0] + m[4] + m[8];
	if (t > 0.0) {
		t = sqrt(t + 1.0);
		w = 0.5 * t;
		t = 0.5 / t;
		x = (m[7] - m[5]) * t;
		y = (m[2] - m[6]) * t;
		z = (m[3] - m[1]) * t;
	} else {
		int i = 0;
		if (m[4] > m[0]) i = 1;
		if (m[8] > m[4 * i]) i = 2;
		const int j = (i + 1) % 3, k = (j + 1) % 3;
		t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
		double v[3];
		v[i] = 0.5 * t;
		t = 0.5 / t;
		w = (m[3 * k + j] - m[3 * j + k]) * t;
		v[j] = (m[3 * j + i] + m[3 * i + j]) * t;
		v[k] = (m[3 * k + i] + m[3 * i + k]) * t;
		x = v[0];
		y = v[1];
		z = v[2];
	}
	double n = sqrt(x * x + y * y + z * z);
	if (n != 0.0) {
		const double angle = 2.0 * atan2(n, fabs(w));
		if (w < 0.0) n = -n;
		rv[0] = angle * (x / n);
		rv[1] = angle * (y / n);
		rv[2] = angle * (z / n);
	} else {
		rv[0] = rv[1] = rv[2] = 0.0;  // angle 0 about (1, 0, 0)
	}
}
// Eigen::AngleAxisd(angle, axis).toRotationMatrix()
OTG_HD inline void angle_axis_matrix(double angle, const double* ax, double* res) {
	const double s = sin(angle), c = cos(angle);
	const double c1[3] = {(1.0 - c) * ax[0], (1.0 - c) * ax[1], (1.0 - c) * ax[2]};
	double tmp;
	tmp = c1[0] * ax[1];
	res[1] = tmp - s * ax[2];
	res[3] = tmp + s * ax[2];
	tmp = c1[0] * ax[2];
	res[2] = tmp + s * ax[1];
	res[6] = tmp - s * ax[1];
	tmp = c1[1] * ax[2];
	res[5] = tmp - s * ax[0];
	res[7] = tmp + s * ax[0];
	res[0] = c1[0] * ax[0] + c;
	res[4] = c1[1] * ax[1] + c;
	res[8] = c1[2] * ax[2] + c;
}

// src/helper_modules/OTG_6dof_cartesian.cpp: position + orientation (rotation vector w.r.t. a reference frame that moves to the
// current orientation whenever the orientation goal changes) through one six-dimensional, phase-synchronised generator
struct CartesianOtg {
	double ref[9];				   // _reference_frame
	double goal_ori[9], goal_w[3];  // _goal_orientation_in_base_frame, _goal_angular_velocity_in_base_frame
	JointsOtg<6> g;				   // _input / _output / trajectory

	OTG_HD void next_orientation(double* R) const {	 // getNextOrientation (:224-235)
		const double* rv = g.out_pos + 3;
		const double n = sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
		double nx[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
		if (!(n < 1e-3)) {
			const double ax[3] = {rv[0] / n, rv[1] / n, rv[2] / n};
			angle_axis_matrix(n, ax, nx);
		}
		m3_mul(ref, nx, R);
	}
	// setGoalPositionAndLinearVelocity (:137-146); NB the reference compares with relative precision 1e-3
	OTG_HD void set_goal_linear(const double* gp, const double* gv) {
		if (is_approx(gp, g.target_pos, 3, 1e-3) && is_approx(gv, g.target_vel, 3, 1e-3)) return;
		g.flags &= ~OTG_GOAL_REACHED;
		g.flags |= OTG_DIRTY;
		for (int i = 0; i < 3; i++) {
			g.target_pos[i] = gp[i];
			g.target_vel[i] = gv[i];
		}
	}
	// setGoalOrientationAndAngularVelocity (:148-188)
	OTG_HD void set_goal_angular(const double* R, const double* w) {
		if (is_approx(goal_ori, R, 9, 1e-3) && is_approx(goal_w, w, 3, 1e-3)) return;
		g.flags &= ~OTG_GOAL_REACHED;
		g.flags |= OTG_DIRTY;
		double new_ref[9], R_new_to_prev[9];
		next_orientation(new_ref);
		m3t_mul(new_ref, ref, R_new_to_prev);
		for (int i = 0; i < 9; i++) {
			ref[i] = new_ref[i];
			goal_ori[i] = R[i];
		}
		for (int i = 0; i < 3; i++) goal_w[i] = w[i];
		double v2[3], a2[3];
		m3_vec(R_new_to_prev, g.out_vel + 3, v2);
		m3_vec(R_new_to_prev, g.out_acc + 3, a2);
		for (int i = 0; i < 3; i++) {
			g.out_pos[3 + i] = 0.0;
			g.out_vel[3 + i] = v2[i];
			g.out_acc[3 + i] = a2[i];
		}
		for (int i = 0; i < 6; i++) {  // _output.pass_to_input(_input): all six components
			g.in_pos[i] = g.out_pos[i];
			g.in_vel[i] = g.out_vel[i];
			g.in_acc[i] = g.out_acc[i];
		}
		double ref_to_goal[9];
		m3t_mul(ref, goal_ori, ref_to_goal);
		rotation_vector(ref_to_goal, g.target_pos + 3);
		m3t_vec(ref, goal_w, g.target_vel + 3);
	}
	OTG_HD void reinitialize_linear(const double* pos) {  // :64-74
		const double zero[3] = {0, 0, 0};
		set_goal_linear(pos, zero);
		for (int i = 0; i < 3; i++) {
			g.in_pos[i] = g.out_pos[i] = g.target_pos[i];
			g.in_vel[i] = g.out_vel[i] = 0.0;
			g.in_acc[i] = g.out_acc[i] = 0.0;
		}
		g.flags |= OTG_DIRTY;
	}
	OTG_HD void reinitialize_angular(const double* R) {	 // :76-87
		const double zero[3] = {0, 0, 0};
		set_goal_angular(R, zero);
		for (int i = 3; i < 6; i++) {
			g.in_pos[i] = g.out_pos[i] = g.target_pos[i];
			g.in_vel[i] = g.out_vel[i] = 0.0;
			g.in_acc[i] = g.out_acc[i] = 0.0;
		}
		g.flags |= OTG_DIRTY;
	}
	OTG_HD void reinitialize(const double* pos, const double* R) {	// :50-62
		const double zero[3] = {0, 0, 0};
		set_goal_linear(pos, zero);
		set_goal_angular(R, zero);
		for (int i = 0; i < 6; i++) {
			g.in_pos[i] = g.out_pos[i] = g.target_pos[i];
			g.in_vel[i] = g.out_vel[i] = 0.0;
			g.in_acc[i] = g.out_acc[i] = 0.0;
		}
		g.flags |= OTG_DIRTY;
	}
	// constructor (:32-48)
	OTG_HD void construct(const double* pos, const double* R) {
		g.flags = 0;
		g.time = g.duration = 0.0;
		for (int i = 0; i < 6; i++) g.target_pos[i] = g.target_vel[i] = g.in_pos[i] = g.in_vel[i] = g.in_acc[i] = g.out_pos[i] = g.out_vel[i] = g.out_acc[i] = 0.0;
		for (int i = 0; i < 9; i++) {
			ref[i] = R[i];
			goal_ori[i] = 0.0;	// uninitialised in the reference; any value that is not a rotation makes the first goal register
		}
		goal_w[0] = goal_w[1] = goal_w[2] = 0.0;
		reinitialize(pos, R);
	}
	// update (:190-222)
	OTG_HD bool update(double dt, const double* vmax6, const double* amax6, Calculator<6>* calc) {
		if (g.flags & OTG_GOAL_REACHED) return false;
		const bool rec = g.update(6, dt, vmax6, amax6, calc);
		if (g.flags & OTG_BAD_FINISH) {	 // finished with a residual velocity: re-target the same pose with zero velocity (:203-206)
			g.flags &= ~OTG_BAD_FINISH;
			const double zero[3] = {0, 0, 0};
			double gp[3] = {g.target_pos[0], g.target_pos[1], g.target_pos[2]};
			set_goal_linear(gp, zero);
			double Rg[9];
			for (int i = 0; i < 9; i++) Rg[i] = goal_ori[i];
			set_goal_angular(Rg, zero);
		}
		return rec;
	}
	// getters (:  OTG_6dof_cartesian.h getNext*)
	OTG_HD void desired(double* pos, double* R, double* v, double* w, double* a, double* al) const {
		for (int i = 0; i < 3; i++) {
			pos[i] = g.out_pos[i];
			v[i] = g.out_vel[i];
			a[i] = g.out_acc[i];
		}
		next_orientation(R);
		m3_vec(ref, g.out_vel + 3, w);
		m3_vec(ref, g.out_acc + 3, al);
	}
};

}  // namespace otg
