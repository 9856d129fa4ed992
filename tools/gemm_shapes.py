"""Histogram of the GEMM shapes of one training step (every ha2g_gemm call, from Python or from inside a launcher).
    python tools/gemm_shapes.py 2> /tmp/gemm.log ; sort /tmp/gemm.log | uniq -c | sort -rn"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ha2g_b200 import graph_step
from ha2g_b200._lib import lib
from ha2g_b200.synthetic import make_batch
from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive
graph_step.enable(False)
dev = torch.device("cuda", 0)
args, gens, D, A, T, (gopts, dopt, aopt, topt) = bench.build_world("expressive", dev)
b = {k: v.to(dev) for k, v in make_batch("expressive", 128, bench.N_WORDS, bench.N_SPEAKERS, seed=1).items()}
call = lambda: train_iter_hierarchy_expressive(args, 11, b["in_text_padded"], b["in_spec"], b["target"], b["vid"], *gens, D, A, T, *gopts, dopt, aopt, topt)
call(); torch.cuda.synchronize()
lib.ha2g_set_gemm_log(1)
call(); torch.cuda.synchronize()
lib.ha2g_set_gemm_log(0)
