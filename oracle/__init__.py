"""CPU oracle for the batched OSC hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: see oracle/primitives.py and DESIGN.md.  Nothing under
sai_primitives_b200/ imports this package.
"""
