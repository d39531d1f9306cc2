#!/usr/bin/env python
"""Development tool: time the K0 lifting kernel on the cfg4 shape."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
spa = importlib.import_module("3dspa_code_b200")
print(json.dumps(bench.run_lifting_leg(spa, torch.device("cuda"))))
