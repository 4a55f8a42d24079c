"""CPU oracle for the batched OSC hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED, except POPCExplicitForceControl, which is checked against the reference's own
source compiled in place (oracle/Makefile -> oracle/_ref, tests/test_popc_reference.py): see oracle/primitives.py and
DESIGN.md section 3.  Nothing under sai_primitives_b200/ imports this package.
"""
