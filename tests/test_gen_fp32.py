"""Build-time transform behind the optional FP32 mode (csrc/gen_fp32.py): the kernel sources rewritten for single precision and
the single-precision mirrors of the parameter structs with their converters.  CPU-only checks: the rewrite rules, and a host
build of the generated mirrors converting a filled OscProgram member by member."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "sai_primitives_b200", "csrc")
sys.path.insert(0, CSRC)
import gen_fp32  # noqa: E402


def test_rewrite_rules():
    src = """namespace osc {
// comment with double and 1.0 stays
DEVI double f(const gdouble* st, double x) { return x * 0.5 + 1e-18 + 2. + fma(x, 3.0e+2, 1) + st[0]; }
#ifndef OSC_FP32_BUILD
DEVI double lean(double d) { return d * 1.5; }
#else
DEVI float lean(float d) { return d * 1.5f; }
#endif  // OSC_FP32_BUILD
asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
int k = 10; double a[3] = {1.0, 2.0, 3.0};
}  // namespace osc"""
    out = gen_fp32.transform(src)
    assert "namespace osc32 {\nusing namespace osc;" in out
    assert "DEVI float f(const gdouble* st, float x) { return x * 0.5f + 1e-18f + 2.f + fma(x, 3.0e+2f, 1) + st[0]; }" in out
    assert "// comment with double and 1.0 stays" in out
    assert "DEVI double lean(double d) { return d * 1.5; }" in out and "DEVI float lean(float d) { return d * 1.5f; }" in out      # verbatim region
    assert 'asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));' in out
    assert "int k = 10; float a[3] = {1.0f, 2.0f, 3.0f};" in out


def test_generated_sources_have_no_double_arithmetic_left(tmp_path):
    gen_fp32.main.__globals__["sys"].argv = ["gen_fp32.py", CSRC, str(tmp_path)]
    gen_fp32.main()
    for f in gen_fp32.FILES:
        text = open(tmp_path / f).read()
        keep = False
        for line in text.split("\n"):
            s = line.strip()
            if s.startswith("#ifdef OSC_FP32_BUILD") or s.startswith("#ifndef OSC_FP32_BUILD"):
                keep = True
            code = line.partition("//")[0]
            if not keep and "asm" not in code:
                assert not re.search(r"\bdouble\b", code), (f, line)
                assert not gen_fp32.LITERAL.search(re.sub(r"(?<![\w.])(?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?f|\d+[eE][-+]?\d+f", "", code)), (f, line)
            if s.startswith("#endif") and "OSC_FP32_BUILD" in s:
                keep = False
    mirrors = open(tmp_path / "osc_types32.h").read()
    assert "struct OscProgram {" in mirrors and "float kp_pos[3];" in mirrors and "const double* q;" in mirrors


def test_mirror_converters_on_the_host(tmp_path):
    gen_fp32.main.__globals__["sys"].argv = ["gen_fp32.py", CSRC, str(tmp_path / "fp32")]
    gen_fp32.main()
    probe = tmp_path / "probe.cpp"
    probe.write_text("""
#include "osc_dev_types.h"
#include "fp32/osc_types32.h"
#include <cstdio>
extern "C" int probe(double* out) {
	static OscProgram P;      // zero-initialised
	double* raw = (double*)&P.model.R_fix[0][0];
	for (int k = 0; k < 9 * OSC_MAX_DOF; k++) raw[k] = 0.1 * k + 1.0 / 3.0;
	P.model.n = 7; P.mft[1].p.kp_pos[2] = 1.0 / 7.0; P.mft[1].p.buffer_size = 200; P.jt[0].p.kv[5] = 2.0 / 3.0; P.n_robots = (1ll << 33) + 5;
	P.mft[0].dt = 1e-3; P.q = (const double*)0x1234560; P.epoch = 77u; P.precision_fp32 = 1; P.model.gravity[2] = -9.81;
	static osc32::OscProgram F;
	osc32::to_f32(P, F);
	int bad = 0;
	for (int j = 0; j < OSC_MAX_DOF; j++)
		for (int k = 0; k < 9; k++) bad += (F.model.R_fix[j][k] != (float)P.model.R_fix[j][k]);
	bad += (F.model.n != 7) + (F.mft[1].p.kp_pos[2] != (float)(1.0 / 7.0)) + (F.mft[1].p.buffer_size != 200) + (F.jt[0].p.kv[5] != (float)(2.0 / 3.0));
	bad += (F.n_robots != (1ll << 33) + 5) + (F.mft[0].dt != 1e-3f) + (F.q != (const double*)0x1234560) + (F.epoch != 77u) + (F.precision_fp32 != 1);
	bad += (F.model.gravity[2] != -9.81f);
	out[0] = (double)sizeof(OscProgram); out[1] = (double)sizeof(osc32::OscProgram);
	return bad;
}
""")
    lib = tmp_path / "libprobe.so"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", CSRC, "-I", str(tmp_path), "-o", str(lib), str(probe)])
    out = (C.c_double * 2)()
    assert C.CDLL(str(lib)).probe(out) == 0
    assert out[1] < out[0]      # the mirror is the smaller struct: its doubles became floats
