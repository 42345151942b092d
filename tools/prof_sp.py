"""Profiling target: two superpixel-graph steps (BASELINE configs[2] shape) for ncu launch lists."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.sweep as sw
print(sw.superpixel_point(B=8, T=8, SP=196)["ms"])
