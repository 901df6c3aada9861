mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_target_r1.json 2> gpurun_out/bench_target_r1.err; cat gpurun_out/bench_target_r1.json; tail -2 gpurun_out/bench_target_r1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_target_ref_r1.json 2>/dev/null; cat gpurun_out/bench_target_ref_r1.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_target_r1.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fir_fft_kernel -s 3 -c 1 -o gpurun_out/prof_target_fir_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 6 -c 1 -o gpurun_out/prof_target_fused_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu1.log gpurun_out/ncu2.log
nproc; lscpu | grep "Model name"
