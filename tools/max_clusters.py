"""How many 2-CTA clusters of the cluster step kernel does this GPU hold at once (cfg2 shapes)?  The in-kernel
gradient reduction needs the whole grid co-resident (B <= that number).  Usage: python tools/max_clusters.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeprank_gnn_b200 import _lib, ops
lib=_lib.load()
sm=ops.ginet_step2_smem_bytes(32,16,32,200,64,32,1000,128,1)
print('smem',sm,'max clusters',lib.drgnn_ginet_step2_max_clusters(sm), lib.drgnn_last_error())
