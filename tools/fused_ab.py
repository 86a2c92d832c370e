#!/usr/bin/env python
"""cifar10 training step with the three PDE branches on side streams (default) against one launch per pass
(`MultiScaleExtractor.fused_branches`), alternated inside one process: ms per step of every run."""
import sys, os
sys.path.insert(0, os.getcwd())
import cnn_with_pde_b200.train as T
import cnn_with_pde_b200.classifiers as Cl
res = {"streams": [], "fused": []}
for rep in range(4):
    for mode in ("streams", "fused"):
        Cl.MultiScaleExtractor.fused_branches = (mode == "fused")
        out = T.run("cifar10", 512, 300, 20, graph=True, quiet=True)
        res[mode].append(round(out["ms_per_step"], 4))
print(res)
