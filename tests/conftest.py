import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    from sai_primitives_b200 import capi
    return capi.load_library()


@pytest.fixture(autouse=True)
def baseline_configs_run_with_internal_otg_off():
    """BASELINE.json's configs run with the tasks' internal OTG off (the reference's examples call disableInternalOtg(),
    examples/01-joint_control/01-joint_control.cpp:136), while the mirrors keep the reference's default (ON, JointTask.h:38,
    MotionForceTask.h:67).  The parity tests of the control law therefore flip the mirror's DefaultParameters, exactly like
    the checkers do (OracleBatch.add_*(otg=False)); the OTG tests (tests/test_gpu_otg.py) put it back."""
    try:
        from sai_primitives_b200 import batched
    except Exception:
        yield
        return
    old = (batched.JointTask.DefaultParameters.use_internal_otg, batched.MotionForceTask.DefaultParameters.use_internal_otg)
    batched.JointTask.DefaultParameters.use_internal_otg = False
    batched.MotionForceTask.DefaultParameters.use_internal_otg = False
    yield
    batched.JointTask.DefaultParameters.use_internal_otg, batched.MotionForceTask.DefaultParameters.use_internal_otg = old
