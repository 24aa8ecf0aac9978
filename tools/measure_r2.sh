# Round-2 ncu evidence (one GPU): the launch list of one step (time only) and a --set full capture of the forward / data-gradient
# convolution instances (fused-pool and plain) + the BatchNorm kernels; raw CSV pages only (.ncu-rep files are too large to bring back).
tag=${1:-r2h}
set -x
timeout 300 python tools/profile_step.py 1024 2 1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py 1024 2 1 > gpurun_out/${tag}_ncu_l.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|bn_relu_apply8|bn_relu_pool8_fwd" -c 26 -o /tmp/${tag}_step python tools/profile_step.py 1024 0 1 > gpurun_out/${tag}_ncu_s.log 2>&1
ncu -i /tmp/${tag}_step.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_conv_raw.csv 2>/dev/null
# source-level view of the fused-pool first audio layer (stall reasons per line): launch ids are resolved by name below
ncu -i /tmp/${tag}_step.ncu-rep --page source --csv --kernel-name regex:"conv_tc_kernel.*TcCfg<1, 8, 16, 112" > gpurun_out/${tag}_ncu_source_A0.csv 2>/dev/null
ls -la gpurun_out/${tag}_*
