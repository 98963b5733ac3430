"""CPU oracle for the hot path -- TEST INFRASTRUCTURE, never imported by the product package.

Each module restates one reference file (cited line by line) with torch-CPU / NumPy primitives:
  temporal_model.py  <- common/models/TemporalModel.py
  camera.py          <- common/camera.py + common/quaternion.py
  loss.py            <- common/loss.py
Parity pin: tests/golden/*.npz, generated from the imported reference by tests/golden/make_golden.py.
Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and --impl reference only).
"""
