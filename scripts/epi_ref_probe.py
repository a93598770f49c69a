import os, sys, glob
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colosseum_b200.hardness as hd
for f in sorted(glob.glob("tests/golden/inst_*_epi.npz")):
    g = np.load(f)
    ref = float(g["diameter"])
    fp = hd.get_diameter(g["T_epi"], True)
    e32 = hd.get_diameter(g["T_epi"], True, precision="f32", epsilon=1e-3)
    e64 = hd.get_diameter(g["T_epi"], True, precision="f64", epsilon=1e-3)
    print(f"{os.path.basename(f):34s} ref {ref:10.5f}  fixed-point {fp:10.5f} ({fp-ref:+.2e})  f32/1e-3 {e32:10.5f} ({e32-ref:+.2e})  f64/1e-3 {e64:10.5f} ({e64-ref:+.2e})")
