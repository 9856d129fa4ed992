#!/bin/bash
# One `ncu --set full` capture per hot kernel of the training step (1 GPU; each capture replays the kernel ~40 times).
# Usage under gpurun:  bash tools/ncu_full.sh <tag>    -> gpurun_out/<tag>_<kernel>.ncu-rep
tag=${1:-r01}
run() {  # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/${tag}_$1 \
      python tools/profile_step.py --steps 1 > gpurun_out/${tag}_$1.log 2>&1
}
run gru_fwd gru_seq_fwd_tc2 40
run gru_bwd gru_seq_bwd_tc2 10
run gemm gemm_packed_kernel 300
run wgrad conv_wgrad_tc2 4
ls -la gpurun_out/${tag}_*.ncu-rep
