# ncu source-level profile of the audio augmentation kernel (per-line executed instructions / stall samples), one step at B = 1024
tag=${1:-r2z}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aug_apply_kernel -s 1 -c 1 -o /tmp/${tag}_aug python tools/profile_step.py 1024 1 1 > gpurun_out/${tag}_ncu_aug.log 2>&1
ncu -i /tmp/${tag}_aug.ncu-rep --page source --csv > gpurun_out/${tag}_aug_source.csv 2>/dev/null
ncu -i /tmp/${tag}_aug.ncu-rep --page raw --csv > gpurun_out/${tag}_aug_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_aug*
