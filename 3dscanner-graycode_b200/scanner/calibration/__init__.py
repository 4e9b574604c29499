"""Calibration (reference scanner/calibration/*) is a one-time offline step outside the
hot path this package accelerates; use the reference's calibrators and feed their .npy
outputs to scanner.triangulation.Triangulate."""
