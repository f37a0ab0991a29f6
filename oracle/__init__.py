"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A restatement, in numpy / plain C, of the arithmetic the reference's hot path performs
(which lives in OpenCV, called through the `opencv` crate 0.88.8 — see DESIGN.md §Oracle).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it; the product (cubesat-apds_b200/) never does.
"""
