#!/usr/bin/env python
"""Launcher entry point: see cnn-with-pde_b200/train.py (the package directory name is not
importable as written, so this file imports it through the cnn_with_pde_b200 shim)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from cnn_with_pde_b200.train import main  # noqa: E402

if __name__ == "__main__":
    main()
