"""FP32 issue ceilings measured with the library's own microbenchmarks (GPU only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sphere_roaming_b200 import engine
from sphere_roaming_b200.forcetable import synthetic_table
from sphere_roaming_b200.types import make_config
with engine.Engine(make_config(), synthetic_table()) as eng:
    a, mhz = eng.fp32_peak()
    b = eng.fp32_peak_rrr()
    x2 = eng.fp32_peak_x2()
    print("FFMA reg x uniform + reg : %.2f TFLOP/s  (clock attr %.0f MHz)" % (a, mhz))
    print("FFMA reg x reg + reg     : %.2f TFLOP/s  (%.3f of the former)" % (b, b / a))
    print("FFMA2 packed, 3 registers : %.2f TFLOP/s  (%.3f of the first)" % (x2, x2 / a))
