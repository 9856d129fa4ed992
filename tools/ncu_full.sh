#!/bin/bash
# One `ncu --set full` capture per hot kernel of the training step (1 GPU; each capture replays the kernel ~40 times).
# Usage under gpurun:  bash tools/ncu_full.sh <tag>    -> gpurun_out/<tag>_<kernel>.ncu-rep
tag=${1:-r02}
run() {  # name regex invocation(0-based among matches);  ONLY="a b" restricts the run to those names
  if [ -n "$ONLY" ] && ! echo " $ONLY " | grep -q " $1 "; then return; fi
  timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/${tag}_$1 \
      python tools/profile_step.py --steps 1 > gpurun_out/${tag}_$1.log 2>&1
}
run gru_fwd28x2 "gru_seq_fwd_tc2_kernel<.int.28, .int.2>" 10
run gru_bwd20 "gru_seq_bwd_tc2_kernel<.int.20>" 10
run gru_fwd10x2 "gru_seq_fwd_tc2_kernel<.int.10, .int.2>" 2
run gemm256 "gemm_packed_kernel<.int.256" 40
run gemm128 "gemm_packed_kernel<.int.128" 40
run pack_pair pack_pair_kernel 80
run splitk splitk_reduce_kernel 40
run conv128 "conv_tc_persist_kernel<.int.128" 10
run conv32 "conv_tc_persist_kernel<.int.32" 4
run wgrad conv_wgrad_tc2 4
run pack_nhwc pack_nhwc_kernel 10
run bn_reduce "bn_reduce_vec_kernel<.int.0>" 3
run bn_apply bn_apply_vec_kernel 3
run bn_bwd bn_bwd_apply_vec_kernel 3
run se_reduce "se_reduce_vec_kernel<.int.1>" 1
run adam adam_multi_dev_kernel 2
run contrastive_rows contrastive_rows_kernel 0
run contrastive_coef contrastive_coef_kernel 0
if [ -z "$ONLY" ] || echo " $ONLY " | grep -q " mel "; then
timeout 200 ncu --set full --clock-control none --kernel-name-base demangled -k regex:mel_power_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_mel \
    python -c "
import sys; sys.path.insert(0, '.')
import torch
from ha2g_b200 import mel
y = torch.randn(9600000, device='cuda') * 0.1
for _ in range(4): mel.extract_melspectrogram(y)
torch.cuda.synchronize()" > gpurun_out/${tag}_mel.log 2>&1
fi
ls -la gpurun_out/${tag}_*.ncu-rep | wc -l
