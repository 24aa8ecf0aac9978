#!/bin/bash
# compute-sanitizer over the tensor-core kernel tests (hand-rolled mbarrier / TMEM / TMA protocols, in-place smem transforms).
# usage: tools/sanitize.sh <out-prefix>      (on a GPU box; logs go to <out-prefix>_{memcheck,racecheck}.log)
out=${1:-gpurun_out/sanitizer}
SEL='first_layer_fused or conv_tc_forward or conv_tc_data_gradient or conv_tc_weight_gradient or linear_tensor_core or prep_weights'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 20 \
    python -m pytest tests/test_conv_tc_gpu.py -q -x -k "$SEL" > ${out}_memcheck.log 2>&1
echo "memcheck rc=$?" >> ${out}_memcheck.log
SEL2='first_layer_fused_backward_matches_apply_then_wgrad or conv_tc_forward'
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 --print-limit 20 \
    python -m pytest tests/test_conv_tc_gpu.py -q -x -k "$SEL2" > ${out}_racecheck.log 2>&1
echo "racecheck rc=$?" >> ${out}_racecheck.log
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed\|rc=" ${out}_memcheck.log ${out}_racecheck.log
