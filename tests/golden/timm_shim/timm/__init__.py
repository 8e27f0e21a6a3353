"""Fake `timm` used ONLY by tests/golden/make_golden.py to import the read-only reference
(`/root/reference/models/ESMStereo.py:11` does `import timm`).  Delegates to the stand-in backbone."""
from esmstereo_b200.backbone import create_model  # noqa: F401

__esm_b200_shim__ = True
