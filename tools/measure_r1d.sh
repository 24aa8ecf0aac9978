# Single-GPU measurement set of a revision: bench (ours + CPU reference arm), other configurations, ncu launch list and one
# ncu --set full capture of the convolution kernels (raw CSV only: .ncu-rep files are too large to bring back).
set -x
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/r1d_bench_b1024.json 2> gpurun_out/r1d_bench_b1024.err; tail -c 400 gpurun_out/r1d_bench_b1024.json
cp gpurun_out/bench_ops_n1_b1024.json gpurun_out/r1d_bench_ops_b1024.json
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1d_bench_reference_cpu.json 2>/dev/null; tail -c 300 gpurun_out/r1d_bench_reference_cpu.json
for cfg in "--batch 128" "--batch 256" "--batch 2048" "--mode mse" "--mode infonce" "--mode infonce --batch 8192" "--mode semi_supervised --batch 2048" "--kind image_simple --batch 256" "--kind image_simple --batch 2048"; do
  n=$(echo $cfg | tr -d ' -' ); timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $cfg > gpurun_out/r1d_other_$n.json 2>/dev/null; cut -c1-160 gpurun_out/r1d_other_$n.json | tail -1
done
timeout 300 python tools/profile_step.py 1024 2 1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1d_launches.csv python tools/profile_step.py 1024 2 1 > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_(kernel|wgrad)" -c 30 -o /tmp/step python tools/profile_step.py 1024 0 1 > gpurun_out/ncu_s.log 2>&1; ncu -i /tmp/step.ncu-rep --page raw --csv > gpurun_out/r1d_ncu_step_raw.csv 2>/dev/null; ls -la gpurun_out/r1d_ncu_step_raw.csv gpurun_out/r1d_launches.csv
