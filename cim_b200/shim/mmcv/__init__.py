"""Import shim: makes `from mmcv.ops import RoIPool, RoIAlign, ...` (lib/ops/__init__.py:6 of the
reference) resolve to cim_b200.ops when cim_b200/shim is on PYTHONPATH.  Only the six names the
reference imports exist."""
__version__ = "1.7.0+cim_b200.shim"
