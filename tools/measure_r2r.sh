# Round-2 ncu evidence for the simple encoder family (one GPU): launch list of one multi_simple step + a --set full capture of the
# K-chunked / N-split convolutions, the per-tap weight gradients and the split first-layer backward.  Raw CSV pages only.
tag=${1:-r2r}
B=${2:-256}
set -x
timeout 300 python tools/profile_step.py $B 2 1 multi_simple || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py $B 2 1 multi_simple > gpurun_out/${tag}_ncu_l.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv_tc_wgrad" -c 40 -o /tmp/${tag}_step python tools/profile_step.py $B 0 1 multi_simple > gpurun_out/${tag}_ncu_s.log 2>&1
ncu -i /tmp/${tag}_step.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_conv_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_*
