"""Runs the emulated NUFFT tests (tests/test_nufft_emulated.py) against a SANITIZER build of the CUDA-on-CPU
driver, see tools/nufft_sanitizers.sh.  argv[1] = path of the instrumented libnufft_emu_*.so."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import test_nufft_emulated as T
class MP:
    def setenv(self, k, v): os.environ[k] = v
lib = ctypes.CDLL(sys.argv[1])
c_vp, c_i64, c_int, c_dbl = T.c_vp, T.c_i64, T.c_int, T.c_dbl
shared = [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_int, c_vp, c_i64, c_dbl, c_dbl, c_vp, c_vp, c_i64, c_int, c_dbl, c_vp]
lib.emu_nufft_shared.argtypes = shared
lib.emu_nufft_shared_chunked.argtypes = shared + [c_int]
lib.emu_nufft_ragged.argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_int, c_vp, c_vp]
lib.emu_last_error.restype = ctypes.c_char_p
cases = [(3, 5.0, 1, 2, 0, {}), (4, 1.0, 1, 1, 0, {}), (5, 5.0, 3, 2, 2, {}),
         (3, 5.0, 1, 2, 0, {"LKB_NUFFT_FFT": "smem"}),
         (3, 5.0, 2, 1, 2, {"LKB_NUFFT_FFT": "smem", "LKB_NUFFT_TWIDDLE_CHAIN": "1", "F": "1100"}),
         (5, 1.0, 1, 2, 0, {"LKB_NUFFT_FFT": "fused", "F": "1100", "LKB_NUFFT_TILE": "1024"})]
for c in cases:
    for k in ("LKB_NUFFT_FFT", "LKB_NUFFT_TWIDDLE_CHAIN", "LKB_NUFFT_TILE", "LKB_NUFFT_GROUP_MB", "LKB_NUFFT_W"):
        os.environ.pop(k, None)
    T.test_shared_grid_translation_unit_on_the_emulator(lib, MP(), *c)
    print("shared ok", c, flush=True)
for fft in ("", "smem"):
    os.environ.pop("LKB_NUFFT_FFT", None)
    T.test_ragged_translation_unit_on_the_emulator(lib, MP(), fft)
    print("ragged ok", repr(fft), flush=True)
os.environ.pop("LKB_NUFFT_FFT", None)
T.test_shared_grid_refuses_unsorted_times(lib)
print("ALL TESTS RAN TO THE END UNDER THE SANITIZER (any report is printed above)")
