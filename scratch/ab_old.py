"""Time the gwc stem / agg / 24->24 layers with the library of the session start (scratch/old_csrc) -- diagnostic."""
import os, sys
sys.path.insert(0, ".")
from esmstereo_b200 import _lib
_lib.LIB_PATH = os.path.abspath("scratch/old_csrc/esmstereo_b200/csrc/libesm_old.so")
_lib.SIGNATURES.pop("esm_set_pdl", None)
sys.argv = ["prof_tc.py", "stem", "agg", "c24"]
import runpy
runpy.run_path("scripts/prof_tc.py", run_name="__main__")
