"""GPU diagnostic: bandwidth h of the tcgen05 path vs the exact SIMT arm vs float64, wide particles (D = 11201)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ocbnn-public_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import ocbnn_oracle as orc
from ocbnn_b200 import engine as E
from helpers import load_golden

def run(X, tag):
    Xd = torch.tensor(X).cuda().contiguous()
    n, D = X.shape
    ws1 = E.svgd_workspace(n, D, Xd.device, use_tensor=1)
    ws0 = E.svgd_workspace(n, D, Xd.device, use_tensor=0)
    for mode in (1, 0):
        E.svgd_select_mode(mode)
        ht = float(E.svgd_bandwidth(Xd, ws1, use_tensor=1))
        print(tag, "tensor mode", mode, ht)
    E.svgd_select_mode(1)
    hs = float(E.svgd_bandwidth(Xd, ws0, use_tensor=0))
    X64 = X.astype(np.float64)
    d = []
    for i in range(n):
        for j in range(i + 1, n):
            d.append(np.sqrt(((X64[i] - X64[j] + 1e-6) ** 2).sum()))
    d = np.sort(np.array(d)); med = d[(len(d) - 1) // 2]
    h64 = med * med / np.log(n)
    print(tag, "simt", hs, "f64", h64, "oracle32", float(orc.svgd_rbf_h(X)), "rel tensor", abs(ht - h64) / h64, "rel simt", abs(hs - h64) / h64)
    print(" sorted d^2 near median:", (d[(len(d) - 1) // 2 - 1:(len(d) - 1) // 2 + 2] ** 2))

g = load_golden("recid_small")
run(g["svgd_x0"], "recid_small golden")
print("golden h", g["svgd_h"][0])
rng = np.random.default_rng(0)
run((0.3 * rng.standard_normal((1, 11201)) + 0.03 * rng.standard_normal((40, 11201))).astype(np.float32), "synthetic n=40")
run((0.8 * rng.standard_normal((40, 200)) + 0.3).astype(np.float32), "D=200 n=40")
