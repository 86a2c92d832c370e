#!/usr/bin/env python
"""fwd / bwd time of bench layers at their roofline batch:  python tools/quick_layer.py emotion tiny ..."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
peak, _ = bench._peaks()
for name in sys.argv[1:]:
    print(name, json.dumps(bench._quick_layer(name, torch.device("cuda", 0), peak)), flush=True)
