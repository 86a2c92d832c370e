import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
names = sys.argv[1:] or list(bench.LAYERS)
for n in names:
    r = bench._quick_layer(n, torch.device("cuda"), 6551.4)
    print(n, r["fwd_ms"], r["bwd_ms"], r["fwd_hbm_frac"], r["bwd_hbm_frac"], flush=True)
