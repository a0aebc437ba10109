import sys, traceback
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from millieye_b200 import _lib
import test_gpu_chain as t
try:
    t.test_chain_thin_layers_and_fp32_head()
    print("PASS")
except Exception as e:
    traceback.print_exc()
    w = _lib.debug_status()
    print("debug word 0x%016x: tag 0x%x cta %d aux 0x%x" % (w, w >> 32, (w >> 8) & 0xffffff, w & 0xff))
