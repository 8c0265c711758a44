import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tests.test_gpu_parity as T
from multipinn_b200 import _lib
orig = T.full_model
for eng, nm in ((_lib.ENGINE_AUTO, "auto"), (_lib.ENGINE_TCGEN05, "tc"), (_lib.ENGINE_FFMA, "ffma"), (_lib.ENGINE_AUTO, "auto")):
    T.full_model = lambda name, seed=0, kernel_engine=None, e=eng: orig(name, seed, e)
    for rep in range(2):
        try:
            print(nm, rep, end=" ")
            T.test_full_depth_step_vs_fp64_autograd("poisson_2D_1C")
        except AssertionError as ex:
            print("ASSERT", str(ex)[:80])
