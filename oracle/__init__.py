"""CPU oracles for the SegDINO3D lifting / superpoint-pool / mask-logit path.

TEST INFRASTRUCTURE ONLY. Nothing under segdino3d_b200/ imports this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""
